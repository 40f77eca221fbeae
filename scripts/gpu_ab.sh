#!/bin/bash
# Same-box A/B of run-time switches on the full pre-training step: every variant twice, interleaved.
# usage: gpu_ab.sh "VAR1=a VAR2=b" "VAR1=c" ...   (each argument = one environment)
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > /dev/null 2>&1
: > gpurun_out/ab.log
for rep in 1 2; do
  for envs in "$@"; do
    ms=$(env $envs python bench.py --steps 24 --warmup 4 --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))")
    echo "rep$rep [$envs] step_ms e2e_ms: $ms" | tee -a gpurun_out/ab.log
  done
done
