import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlxmert_b200 import params as P
from xlxmert_b200.generator import B200Generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
code = torch.rand(B, 8, 8, 2048, device="cuda") * 0.1
for _ in range(2):
    G(code, train=False)
torch.cuda.synchronize()
