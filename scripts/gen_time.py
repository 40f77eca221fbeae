#!/usr/bin/env python
"""Generator forward timing (B=128 by default), CUDA events, device-resident codes.  Usage: gen_time.py [B]"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
e.build()
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from xlxmert_b200.generator import B200Generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
ids = torch.randint(0, D.num_clusters, (B, 64), device="cuda")
code = synth.centroid_table(D).cuda()[ids]
f = lambda: G(code.view(B, 8, 8, 2048), train=False)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 8
print(json.dumps({"B": B, "ms": ms, "img_per_s": B / ms * 1e3, "algorithmic_tflops": 27.755 * B / ms, "env": {k: v for k, v in os.environ.items() if k.startswith("XLX_")}}))
