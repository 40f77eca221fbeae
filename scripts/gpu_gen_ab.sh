#!/bin/bash
# generator parity + same-box A/B of one environment switch: gpu_gen_ab.sh VAR  (VAR=0 vs VAR=1, interleaved)
mkdir -p gpurun_out
VAR=${1:-XLX_GEMM_TMA_OUT_BIAS}
timeout 900 python -m pytest tests/test_generator_parity.py tests/test_edge_cases.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
  for v in 0 1; do
    env $VAR=$v timeout 300 python scripts/gen_time.py 128 | tail -1
  done
done
