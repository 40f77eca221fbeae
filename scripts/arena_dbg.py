import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_encoder_parity import _case, _encoder, O, TINY_DIMS
sd, batch, feats, emb, mask = _case(TINY_DIMS, 2, 5, 6, 3, 4)
enc = _encoder(TINY_DIMS, O.sub(sd, "encoder")).train()
uc = torch._C._storage_Use_Count
def counts():
    return [(hex(t.data_ptr()), uc(t.untyped_storage()._cdata)) for t in enc.__dict__.get("_arena_pool", [])]
def step():
    e = emb.cuda().requires_grad_(True)
    (v, _), (l, _), _ = enc(e, mask.cuda(), feats.cuda(), batch["visual_pos"].cuda())
    (l[-1].sum() + (v[-1] ** 2).sum()).backward()
for i in range(6):
    if i != 1:
        for p in enc.parameters():
            p.grad = None
    print("before step", i, counts())
    step()
    print("after  step", i, counts(), "arena", hex(enc.last_grad_arena.data_ptr()))
