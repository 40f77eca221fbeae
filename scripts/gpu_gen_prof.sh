#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_gen.csv python scripts/gen_prof.py 32 > gpurun_out/ncu_gen.log 2>&1
echo "exit=$?"
