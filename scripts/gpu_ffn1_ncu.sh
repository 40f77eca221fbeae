#!/bin/bash
# ncu --set full (+ source counters) of one FFN-1 forward GEMM launch through the probe binary:
# M=16384 N=3072 K=768, 3 passes, bias + GeLU + saved gelu' + split output (epi 139)
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > /dev/null 2>&1
SHAPE=${1:-"16384 3072 768 3 0 0 139"}
NAME=${2:-ffn1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -f -o gpurun_out/$NAME ./xlxmert_b200/lib/gemm_test $SHAPE 3 > gpurun_out/$NAME.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/$NAME.log
ncu -i gpurun_out/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i gpurun_out/$NAME.ncu-rep --page source --csv > gpurun_out/${NAME}_source.csv 2>/dev/null
ls -la gpurun_out/$NAME*
