#!/usr/bin/env python
"""Per-launch table of the generator's tcgen05 GEMM / convolution launches at batch B (one forward, events around every
launch — serialised, so the sum exceeds the overlapped forward).  Usage: gen_gemm_table.py [B] [out.csv]
Honours XLX_GEMM_DEBUG (pipeline-ablation bits) and every other XLX_* switch."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
e.build()
from xlxmert_b200 import _lib, params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from xlxmert_b200.generator import B200Generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = sys.argv[2] if len(sys.argv) > 2 else None
if out:
    os.environ["XLX_GEMM_LOG"] = out
lib = _lib.load()
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
ids = torch.randint(0, D.num_clusters, (B, 64), device="cuda")
code = synth.centroid_table(D).cuda()[ids]
for _ in range(2):
    G(code.view(B, 8, 8, 2048), train=False)
torch.cuda.synchronize()
lib.xlx_profile_gemm_begin()
G(code.view(B, 8, 8, 2048), train=False)
tot_ms, tot_fl, n = C.c_double(), C.c_double(), C.c_int64()
_lib.check("xlx_profile_gemm_end", lib.xlx_profile_gemm_end(C.byref(tot_ms), C.byref(tot_fl), C.byref(n)))
print(f"debug={os.environ.get('XLX_GEMM_DEBUG', '0')} launches={n.value} gemm_ms={tot_ms.value:.3f} "
      f"algorithmic_tflops={tot_fl.value / tot_ms.value / 1e9:.1f}")
