#!/bin/bash
mkdir -p gpurun_out
for cfg in "16384 3072 768 3 0 0 0" "16384 3072 768 1 0 0 0" "16384 3072 768 3 0 0 139" "3072 768 16384 3 1 1 0"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_$tag ./build/gemm_test $cfg 3 > gpurun_out/ncu_$tag.log 2>&1
  echo "$cfg exit=$?"
done
ls -la gpurun_out/*.ncu-rep
