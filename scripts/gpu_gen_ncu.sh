#!/bin/bash
# Generator at B=128: ncu --set full of every tcgen05 GEMM / convolution launch of one forward; only the raw CSV page
# comes back (the .ncu-rep of ~45 launches exceeds gpurun's 64 MiB return limit).
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > /dev/null 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:gemm_kernel -s 0 -c 400 -f -o /tmp/gen_full python scripts/gen_once.py 128 1 > gpurun_out/gen_full.log 2>&1
echo "full exit=$?"
ncu -i /tmp/gen_full.ncu-rep --page raw --csv > gpurun_out/gen_full_raw.csv 2>/dev/null
ls -la gpurun_out/gen_full_raw.csv
