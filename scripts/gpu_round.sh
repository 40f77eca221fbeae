#!/bin/bash
# One GPU visit: parity tests, the bench line, the reference arm, an ncu launch list of one step and a full
# ncu capture of the dominant kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
XLX_GEMM_LOG=gpurun_out/gemm_shapes.csv timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 300 -c 3 -f -o gpurun_out/prof_gemm \
      python bench.py --ncu --steps 1 --warmup 0 > gpurun_out/ncu_gemm.log 2>&1
  echo "ncu gemm exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 60 -c 4 -f -o gpurun_out/prof_attn \
      python bench.py --ncu --steps 1 --warmup 0 > gpurun_out/ncu_attn.log 2>&1
  echo "ncu attn exit=$?"
fi
ls -la gpurun_out
