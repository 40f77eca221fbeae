#!/bin/bash
# A/B of the standalone GEMM micro-benchmarks with and without the dependent-launch attribute
mkdir -p gpurun_out
for pdl in 0 1; do
  echo "=== XLX_PDL=$pdl"
  for cfg in "16384 3072 768 3 0 0 0" "16384 3072 768 3 0 0 139" "16384 768 3072 3 0 0 5" "16384 2304 768 3 0 0 8" "5120 3072 768 3 0 0 139" "3072 768 16384 3 1 1 0"; do
    echo "--- $cfg"; XLX_PDL=$pdl timeout 120 ./build/gemm_test $cfg 20 2>&1 | grep "time"
  done
done
