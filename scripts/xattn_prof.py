#!/usr/bin/env python
"""One cross-modality attention block in isolation (for ncu): an encoder with l_layers = r_layers = 0, x_layers = 1 at
the bench batch (B = 256, L = 20, V = 64), forward only (or forward + backward with `bwd`), everything on one stream
(XLX_TWO_STREAMS=0) so that the launch order is the program order: visual feature encoder, then LxmertXLayer =
cross-attention block [QKV projection GEMM over all B·(L+V) rows (shared weights), attention core lang→vis, attention
core vis→lang, output projection GEMM (+bias +residual), LayerNorm], then the two self-attention blocks and FFNs."""
import os, sys
os.environ.setdefault("XLX_TWO_STREAMS", "0")
os.environ.setdefault("XLX_PDL", "0")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
e.build()
from dataclasses import replace
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS
from xlxmert_b200.encoder import B200LxmertEncoder
d = replace(DEFAULT_DIMS, l_layers=0, r_layers=0, x_layers=1)
B = int(os.environ.get("XATTN_B", "256"))
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
sd = P.init_state_dict(P.model_param_specs(d), seed=0)
enc = B200LxmertEncoder(dims=d)
enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
enc = enc.cuda().train()
batch = synth.make_batch(d, B, 20, 64, seed=0)
emb = torch.randn(B, 20, d.hidden).cuda()
feats = synth.centroid_table(d).cuda()[batch["cluster_ids"].cuda()]
mask = ((1.0 - batch["attention_mask"].cuda()[:, None, None, :].float()) * torch.finfo(torch.float32).min)
pos = batch["visual_pos"].cuda()
for it in range(3):
    if bwd:
        e_ = emb.clone().requires_grad_(True)
        (v, _), (l, _), _ = enc(e_, mask, feats, pos)
        (l[-1].sum() + v[-1].sum()).backward()
        for p in enc.parameters():
            p.grad = None
    else:
        with torch.no_grad():
            enc.train()
            (v, _), (l, _), _ = enc(emb.requires_grad_(False), mask, feats, pos)
torch.cuda.synchronize()
print("done")
