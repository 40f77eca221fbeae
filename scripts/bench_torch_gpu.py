#!/usr/bin/env python
"""Context numbers (not the driver's bench): the reference's own encoder class (HF LxmertEncoder, the code
x-lxmert/src/lxrt/modeling.py:5 imports) run eagerly by PyTorch on the same B200 for the bench workload
(B=256, L=20, V=64, fwd+bwd), in fp32 (TF32 off / on) and under bf16 autocast, next to this repo's passes=3 / passes=1."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlxmert_b200 import params as P, synth  # noqa: E402
from xlxmert_b200.config import DEFAULT_DIMS as D  # noqa: E402


def timed(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    from transformers import LxmertConfig
    from transformers.models.lxmert.modeling_lxmert import LxmertEncoder
    B, L, V = 256, 20, 64
    dev = "cuda"
    sd = P.init_state_dict(P.model_param_specs(D), seed=0)
    cfg = LxmertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    enc = LxmertEncoder(cfg)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    enc = enc.to(dev).train()
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(B, L, D.hidden, generator=g).to(dev)
    feats = torch.randn(B, V, D.feat_dim, generator=g).abs().to(dev)
    batch = synth.make_batch(D, B, L, V, seed=0)
    pos = batch["visual_pos"].to(dev)
    mask = ((1.0 - batch["attention_mask"][:, None, None, :].float()) * torch.finfo(torch.float32).min).to(dev)
    gl = (torch.randn(B, L, D.hidden, generator=g) / (B * L)).to(dev)
    gv = (torch.randn(B, V, D.hidden, generator=g) / (B * V)).to(dev)

    def step(autocast=False):
        def f():
            enc.zero_grad(set_to_none=True)
            e = emb.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                (vs, _), (ls, _), _ = enc(e, mask, feats, pos)
            torch.autograd.backward([ls[-1].float(), vs[-1].float()], [gl, gv])
        return f
    out = {}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ms = timed(step())
    out["torch_eager_fp32_tf32off"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = timed(step())
    out["torch_eager_fp32_tf32on"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    ms = timed(step(True))
    out["torch_eager_autocast_bf16"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    del enc
    torch.cuda.empty_cache()

    import __graft_entry__ as entry
    entry.build()
    from xlxmert_b200.encoder import B200LxmertEncoder
    from oracle import lxrt_oracle as O
    for passes in (3, 1):
        mine = B200LxmertEncoder(dims=D, passes=passes)
        mine.load_state_dict(O.sub(sd, "encoder"), strict=True)
        mine = mine.to(dev).train()

        def f():
            mine.invalidate_prepared()
            e = emb.clone().requires_grad_(True)
            (vs, _), (ls, _), _ = mine(e, mask, feats, pos)
            torch.autograd.backward([ls[-1], vs[-1]], [gl, gv])
            for p in mine.parameters():
                p.grad = None
        ms = timed(f)
        out[f"xlxmert_b200_passes{passes}"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
        del mine
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
