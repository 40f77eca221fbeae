#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gemm_exp.log
: > $LOG
run() { echo "--- DEBUG=${XLX_GEMM_DEBUG:-0} $*" >> $LOG; timeout 120 ./build/gemm_test "$@" 2>&1 | grep time >> $LOG; }
for d in 0 1 64 65; do
export XLX_GEMM_DEBUG=$d
run 16384 3072 768 3 0 0 0 20
run 16384 3072 768 3 0 0 139 20
run 16384 768 3072 3 0 0 5 20
run 5120 3072 768 3 0 0 0 20
done
cat $LOG
