#!/bin/bash
# 2-GPU visit: staged-backward parity, then the data-parallel bench with the early language-range reduce on / off,
# interleaved on the same box (run with gpurun --gpus 2).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_parity.py -q -m gpu -k "stage" -x 2>&1 | tail -3
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 2> gpurun_out/bench2.err | tail -1
}
for i in 1 2; do
  run XLX_EARLY_LANGUAGE_REDUCE=0 > gpurun_out/bench2_late_$i.json
  run XLX_EARLY_LANGUAGE_REDUCE=1 > gpurun_out/bench2_early_$i.json
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench2_*.json")):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, r["ms_per_step"], r.get("gradient_exchange"), r["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
P
tail -5 gpurun_out/bench2.err
