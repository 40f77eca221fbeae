#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gemm_exp.log
: > $LOG
export XLX_TEST_MAP=1
run() { echo "--- $*" >> $LOG; timeout 120 ./build/gemm_test "$@" 2>&1 | sed 's/max_abs_err.*rel=/rel=/' | grep -v "mtile" | head -24 >> $LOG; }
for amn in 0 1; do for bmn in 0 1; do
run 512 512 64 3 $amn $bmn 0
run 640 264 96 3 $amn $bmn 13
run 384 768 128 1 $amn $bmn 5
run 1000 776 200 3 $amn $bmn 13
done; done
unset XLX_TEST_MAP
grep -c OK $LOG; grep FAIL $LOG
