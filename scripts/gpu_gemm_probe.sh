#!/bin/bash
# Runs the tcgen05 GEMM probe matrix on the GPU box; each case in its own process under a timeout so a
# trap/hang in one variant does not hide the others.  Output: gpurun_out/gemm_probe.log
mkdir -p gpurun_out
LOG=gpurun_out/gemm_probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
run() { echo "--- XLX_GEMM_BK=${XLX_GEMM_BK:-32} $*" >> $LOG; timeout 60 ./build/gemm_test "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
run 128 256 64 1 0 0 0
run 128 256 768 3 0 0 0
run 256 512 768 3 0 0 15
run 1000 776 200 3 0 0 5
run 256 512 768 3 0 1 0
run 256 512 768 3 1 0 0
run 256 512 768 3 1 1 0
run 1000 776 200 3 1 1 13
run 300 64 768 3 0 0 1
run 300 128 768 1 0 1 1
run 512 768 768 3 0 0 52
export XLX_GEMM_BK=64
run 128 256 768 3 0 0 0
run 1000 776 200 3 0 0 5
run 256 512 768 3 1 1 0
run 1000 776 200 1 1 1 13
unset XLX_GEMM_BK
echo "=== perf" >> $LOG
run 16384 3072 768 3 0 0 0 20
run 16384 768 3072 3 0 0 0 20
run 16384 768 768 3 0 0 0 20
run 16384 3072 768 1 0 0 0 20
run 16384 768 3072 1 0 0 0 20
run 768 3072 16384 3 1 1 0 20
run 16384 3072 768 3 0 1 0 20
export XLX_GEMM_BK=64
run 16384 3072 768 3 0 0 0 20
run 16384 3072 768 1 0 0 0 20
unset XLX_GEMM_BK
tail -80 $LOG
