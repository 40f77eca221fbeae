import os, sys, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlxmert_b200 import params as P, _lib
from xlxmert_b200.generator import B200Generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = _lib.load()
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
code = torch.rand(B, 8, 8, 2048, device="cuda") * 0.1
for _ in range(2):
    G(code, train=False)
torch.cuda.synchronize()
os.environ["XLX_GEMM_LOG"] = "gpurun_out/gen_gemm_shapes.csv"
lib.xlx_profile_gemm_begin()
for _ in range(2):
    G(code, train=False)
a, b, n = C.c_double(), C.c_double(), C.c_int64()
lib.xlx_profile_gemm_end(C.byref(a), C.byref(b), C.byref(n))
print("gemm ms per fwd", a.value / 2, "TFLOP/s", b.value / a.value / 1e9, "launches", n.value // 2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    G(code, train=False)
e1.record(); torch.cuda.synchronize()
print("fwd ms", e0.elapsed_time(e1) / 5, "img/s", B * 5 / e0.elapsed_time(e1) * 1e3)
