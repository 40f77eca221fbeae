"""Diagnostic: does a training step leave reference cycles that keep CUDA tensors alive until the cyclic GC runs?"""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from xlxmert_b200.lxmert import B200LxmertModel

torch.manual_seed(0)
model = B200LxmertModel(D).cuda().train()
B = 64
batch = {k: v.cuda() for k, v in synth.make_batch(D, B, 20, 64, seed=0).items()}
table = torch.randn(D.num_clusters, D.feat_dim, device="cuda")
gc.collect(); gc.disable()
for it in range(8):
    feats = table[batch["cluster_ids"]]
    out = model(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
                attention_mask=batch["attention_mask"])
    loss = out[0].sum() + out[1].sum() + out[2].mean()
    loss.backward()
    torch.cuda.synchronize()
    for p in model.parameters():
        p.grad = None
    del feats, out, loss
    print(it, "allocated MB", round(torch.cuda.memory_allocated() / 2**20, 1), "reserved MB",
          round(torch.cuda.memory_reserved() / 2**20, 1))
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
print("cyclic garbage objects:", n, "allocated after collect MB", round(torch.cuda.memory_allocated() / 2**20, 1))
kinds = {}
for o in gc.garbage:
    kinds[type(o).__name__] = kinds.get(type(o).__name__, 0) + 1
print(sorted(kinds.items(), key=lambda kv: -kv[1])[:15])
