#!/bin/bash
# GPU parity suite on the B200 box; full log (with the arg-max flip counts the tests print) → gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
timeout ${1:-2400} python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^\[argmax|worst parameter|FAILED|Error" gpurun_out/pytest_gpu.log | head -40
