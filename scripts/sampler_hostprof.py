"""Diagnostic: host-side profile of the NAR sampler at B=32 (where the step is launch/host bound)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS as D
from test_pretrain_parity import build_model
from xlxmert_b200.generator import B200Generator
from xlxmert_b200.sampler import B200ImggenModel

G = B200Generator(); G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True); G = G.cuda().eval()
pre, table = build_model(0)
m = B200ImggenModel(D, num_clusters=D.num_clusters)
m.set_visual_embedding(table.clone())
m.load_state_dict({k: v for k, v in pre.state_dict().items() if not k.startswith("cls.")}, strict=False)
m.set_image_generator(G); m = m.cuda()
tok = synth.make_batch(D, 32, 20, 64, seed=3)["input_ids"].cuda()
for _ in range(3):
    m.sample_image_NAR(tok, n_steps=4)
torch.cuda.synchronize()
for name, kw in (("cached", {}), ("no_cache", {"cache_language": False})):
    t0 = time.perf_counter()
    for _ in range(5):
        m.sample_image_NAR(tok, n_steps=4, **kw)
    t1 = time.perf_counter()          # host time to issue (no sync)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(name, "host issue ms/iter", (t1 - t0) / 5 * 1e3, "total ms/iter", (t2 - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    m.sample_image_NAR(tok, n_steps=4)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
