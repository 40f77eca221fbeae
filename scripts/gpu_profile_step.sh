#!/bin/bash
# ncu evidence for the full pre-training step (bench.py --ncu = W warm-up + K resident steps, tasks round-robin):
#  1. launch list (gpu__time_duration.sum) of 3 steps  → per-kernel shares
#  2. DRAM bytes of every tcgen05 GEMM launch of those 3 steps → roofline.traffic
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
python scripts/launch_table.py gpurun_out/launches_step.csv 2 | head -40
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_kernel --csv --log-file gpurun_out/gemm_traffic_step.csv \
    python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit=$?"
