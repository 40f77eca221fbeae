#!/bin/bash
# compute-sanitizer over the paths added late in round 1: halo convolutions, two-stream encoder, row compaction, k-means
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_generator_parity.py -m gpu -x -q -k "golden" > gpurun_out/sanitize_memcheck_generator.log 2>&1
echo "memcheck generator exit=$?"; tail -4 gpurun_out/sanitize_memcheck_generator.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_encoder_parity.py tests/test_kmeans.py tests/test_edge_cases.py -m gpu -x -q -k "tiny or staged or labelled or compaction or 257" > gpurun_out/sanitize_memcheck_late.log 2>&1
echo "memcheck late exit=$?"; tail -4 gpurun_out/sanitize_memcheck_late.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_generator_parity.py -m gpu -x -q -k "golden" > gpurun_out/sanitize_racecheck_generator.log 2>&1
echo "racecheck generator exit=$?"; tail -4 gpurun_out/sanitize_racecheck_generator.log
