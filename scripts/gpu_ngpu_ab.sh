#!/bin/bash
# N-GPU visit (gpurun --gpus N): data-parallel bench with the early language-range reduce off / on, interleaved
N=${1:-8}
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 2> gpurun_out/benchN.err | tail -1
}
for i in 1 2; do
  run XLX_EARLY_LANGUAGE_REDUCE=0 > gpurun_out/bench${N}_late_$i.json
  run XLX_EARLY_LANGUAGE_REDUCE=1 > gpurun_out/bench${N}_early_$i.json
done
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/bench${N}_*.json")):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(r["ms_per_step"], 2), r.get("gradient_exchange", {}).get("exposed_exchange_ms"), round(r["value"]), round(r["e2e"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
P
tail -3 gpurun_out/benchN.err
