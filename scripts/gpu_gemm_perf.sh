#!/bin/bash
# GEMM micro-benchmarks for the shapes of the encoder step (gemm_test M N K passes a_mn b_mn epi reps)
mkdir -p gpurun_out
LOG=gpurun_out/gemm_perf.log
: > $LOG
BIN=./build/gemm_test
run() { echo "--- $BIN BK=${XLX_GEMM_BK:-32} BN=${XLX_GEMM_BN:-auto} $*" >> $LOG; timeout 120 $BIN "$@" 2>&1 | sed 's/max_abs_err.*rel=/rel=/' >> $LOG; }
for BIN in ./build/gemm_test_e4 ./build/gemm_test_e8 ./build/gemm_test_e16; do
  run 16384 3072 768 3 0 0 0 20
  run 16384 3072 768 3 0 0 139 20
  run 16384 3072 768 3 0 1 72 20
  run 16384 768 3072 3 0 0 5 20
  run 16384 768 768 3 0 0 5 20
  run 5120 768 768 3 0 0 5 20
  run 16384 3072 768 1 0 0 0 20
  run 16384 3072 768 1 0 0 8 20
done
cat $LOG
