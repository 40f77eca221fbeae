#!/usr/bin/env python
"""Host-side launch cost vs device time of one pre-training step (diagnostic)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from test_pretrain_parity import build_model

model, table = build_model(0)
model.train()
B = 256
batch = {k: v.cuda() for k, v in synth.make_batch(D, B, 20, 64, seed=0).items()}
labels = dict(word_labels=batch["word_labels"], obj_labels=batch["obj_labels"], matched_labels=batch["matched_labels"])
out = {}
for compact in ((True, False) if os.environ.get("REV") else (False, True)):
    model.obj_predict_head.compact_rows = compact
    model.cls.compact_rows = compact
    for task in ("vis_mask", "word_mask"):
        ids = batch["masked_input_ids"] if task == "word_mask" else batch["input_ids"]
        hf = hb = 0.0
        N = 6
        for it in range(3 + N):
            if it == 3:
                torch.cuda.synchronize(); t_start = time.perf_counter(); hf = hb = 0.0
            for p in model.parameters():
                p.grad = None
            t0 = time.perf_counter()
            o = model(input_ids=ids, visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
                      cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"], label_dict=labels, task=task)
            t1 = time.perf_counter()
            o["total_loss"].backward()
            t2 = time.perf_counter()
            hf += t1 - t0; hb += t2 - t1
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t_start) / N * 1e3
        out[f"{task}_compact{int(compact)}"] = dict(step_ms=round(tot, 2), host_fwd_ms=round(hf / N * 1e3, 2),
                                                    host_bwd_ms=round(hb / N * 1e3, 2))
print(json.dumps(out, indent=1))
