#!/bin/bash
# GEMM correctness matrix + micro-benchmarks for the shapes of the encoder step
# (gemm_test M N K passes a_mn b_mn epi reps; epi bits in gemm_test.cu)
mkdir -p gpurun_out
LOG=gpurun_out/gemm_perf.log
: > $LOG
BIN=./build/gemm_test
run() { echo "--- $*" >> $LOG; timeout 120 $BIN "$@" 2>&1 | sed 's/max_abs_err.*rel=/rel=/' >> $LOG; echo "exit=$?" >> $LOG; }
# correctness: every operand-major combination, ragged shapes, every epilogue
for amn in 0 1; do for bmn in 0 1; do
  run 256 512 768 3 $amn $bmn 0
  run 1000 776 200 3 $amn $bmn 13
  run 1000 776 200 1 $amn $bmn 5
done; done
run 300 64 768 3 0 0 1
run 300 32 288 3 0 0 1
run 512 768 768 3 0 0 52
run 512 768 768 3 0 0 139
run 512 768 768 3 0 1 72
run 768 768 16384 3 1 1 0
echo "=== perf" >> $LOG
run 16384 3072 768 3 0 0 0 20
run 16384 3072 768 3 0 0 139 20
run 16384 3072 768 3 0 1 72 20
run 16384 768 3072 3 0 0 5 20
run 16384 768 3072 3 0 1 4 20
run 16384 768 768 3 0 0 5 20
run 5120 3072 768 3 0 0 139 20
run 5120 768 3072 3 0 0 5 20
run 5120 768 768 3 0 0 5 20
run 3072 768 16384 3 1 1 0 20
run 16384 3072 768 1 0 0 0 20
run 16384 3072 768 1 0 0 8 20
run 16384 768 3072 1 0 0 0 20
grep -c "OK" $LOG; grep "FAIL\|rc=" $LOG; grep -A3 "=== perf" $LOG | head -2; grep "^---\|time" $LOG | sed -n '/16384 3072 768 3 0 0 0 20/,$p'
