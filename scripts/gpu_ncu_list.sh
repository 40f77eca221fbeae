#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
