#!/bin/bash
mkdir -p gpurun_out
for cfg in "16384 3072 768 3 0 0 0"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof2_gemm_$tag ./build/gemm_test $cfg 3 > gpurun_out/ncu_$tag.log 2>&1
  echo "$cfg exit=$?"
done
