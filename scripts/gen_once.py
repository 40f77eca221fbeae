#!/usr/bin/env python
"""N generator forwards at batch B (for ncu: capture the last one).  Usage: gen_once.py [B] [N]; PDL off so that launch
order is program order and no kernel overlaps its predecessor."""
import os, sys
os.environ.setdefault("XLX_PDL", "0")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
e.build()
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from xlxmert_b200.generator import B200Generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
ids = torch.randint(0, D.num_clusters, (B, 64), device="cuda")
code = synth.centroid_table(D).cuda()[ids]
for _ in range(N):
    G(code.view(B, 8, 8, 2048), train=False)
    torch.cuda.synchronize()
