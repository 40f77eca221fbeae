#!/bin/bash
# generator: parity tests, timing, ncu launch list (B=32) → per-kernel shares
mkdir -p gpurun_out
python -m pytest tests/test_generator_parity.py tests/test_edge_cases.py -m gpu -q -x 2>&1 | tail -3
python scripts/gen_time.py 128
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_gen.csv python scripts/gen_time.py 32 > gpurun_out/ncu_gen.log 2>&1
python scripts/launch_table.py gpurun_out/launches_gen.csv 11 | head -24
