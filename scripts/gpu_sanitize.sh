#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_encoder_parity.py -m gpu -x -q -k "tiny or ragged or rejects" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit=$?"; tail -5 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_encoder_parity.py -m gpu -x -q -k "tiny" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit=$?"; tail -5 gpurun_out/sanitize_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_pretrain_parity.py -m gpu -x -q -k "vis_mask or matched" > gpurun_out/sanitize_memcheck_heads.log 2>&1
echo "memcheck heads exit=$?"; tail -5 gpurun_out/sanitize_memcheck_heads.log
