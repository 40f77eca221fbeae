#!/bin/bash
# One GPU-box visit: parity suite, the driver's bench line (both arms), logs under gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_pytest.sh 2400
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_ref.json
