#!/bin/bash
# DRAM traffic of every tcgen05 GEMM launch of one encoder step (for bench.py's roofline.traffic), plus one
# --set full capture of the three dominant GEMM instantiations and the attention kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_kernel -s 321 -c 320 --csv --log-file gpurun_out/gemm_traffic.csv \
    python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 400 -c 6 -f -o gpurun_out/prof_gemm_step \
    python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 40 -c 4 -f -o gpurun_out/prof_attn_step \
    python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
echo "attn full exit=$?"
