#!/usr/bin/env python
"""Where the time of a B=32, 4-step NAR sampling run goes (CUDA events around the phases, synchronised)."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from xlxmert_b200.generator import B200Generator
from xlxmert_b200.sampler import B200ImggenModel
from xlxmert_b200.synth import box_position
dev = torch.device("cuda", 0)
pre, table = bench.build_pretraining_model(dev, 3)
m = B200ImggenModel(D, num_clusters=D.num_clusters)
m.set_visual_embedding(table.clone())
m.load_state_dict({k: v for k, v in pre.state_dict().items() if not k.startswith("cls.")}, strict=False)
G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True)
m.set_image_generator(G); m = m.to(dev).eval(); m.set_visual_embedding(table)
B = 32
tok = synth.make_batch(D, B, 20, 64, seed=3)["input_ids"].to(dev)
vpos = torch.from_numpy(box_position(8)).unsqueeze(0).expand(B, -1, -1).contiguous().to(dev)
def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
out = {}
with torch.no_grad():
    out["language_stack_ms"] = T(lambda: m.bert.language_stack(tok, tok > 0))
    lang = m.bert.language_stack(tok, tok > 0)
    code = torch.zeros(B, 64, 2048, device=dev); vm = torch.empty(B, 64, dtype=torch.uint8, device=dev)
    m._nar_update(code, vm, None, None, 64)
    out["predict_step_ms"] = T(lambda: m._predict(tok, code, vpos, lang))
    out["predict_step_graph_ms"] = T(lambda: m._graph_predict(tok, code, vpos, lang))
    pp, pid = m._predict(tok, code, vpos, lang)
    out["transition_ms"] = T(lambda: m._nar_update(code, vm, pp, pid, 48))
    out["generator_ms"] = T(lambda: m.G(code.view(B, 8, 8, 2048), train=False))
    out["decode_incl_d2h_ms"] = T(lambda: m._decode(code, B, 2048, 8))
    out["full_nar4_ms"] = T(lambda: m.sample_image_NAR(tok, n_steps=4))
    out["full_nar4_graph_ms"] = T(lambda: m.sample_image_NAR(tok, n_steps=4, cuda_graph=True))
    # host-only cost of issuing one predict step (no sync inside)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m._predict(tok, code, vpos, lang)
    out["predict_host_issue_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()
print(json.dumps(out))
