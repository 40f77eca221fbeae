#!/bin/bash
# ncu --set full of every kernel of one isolated cross-modality layer forward (3rd iteration), raw CSV → gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > /dev/null 2>&1
# launch list first (to find the per-iteration launch count), then the full capture of the last iteration
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/xattn_list.csv python scripts/xattn_prof.py > gpurun_out/xattn_list.log 2>&1
n=$(grep -c "gpu__time_duration.sum" gpurun_out/xattn_list.csv); per=$((n / 3)); echo "launches total $n per iteration $per"
timeout 1500 ncu --set full --clock-control none --import-source on -s $((2 * per)) -c $per -f -o gpurun_out/xattn_full python scripts/xattn_prof.py > gpurun_out/xattn_full.log 2>&1
echo "full exit=$?"
ncu -i gpurun_out/xattn_full.ncu-rep --page raw --csv > gpurun_out/xattn_full_raw.csv 2>/dev/null
ls -la gpurun_out/xattn_full* | head
