#!/usr/bin/env python
"""Per-GEMM-class operand-precision study for the encoder (BASELINE.md §2 asks for it before any
use of a reduced-precision MMA kind).  CPU emulation through the oracle (test tooling, never the product path):
every nn.Linear product y = x·Wᵀ of the encoder is computed in fp64 from operands rounded to the format under test,
separately for the FORWARD product, the DGRAD product (dX = dY·W) and the WGRAD product (dW = dYᵀ·X), so one class at
a time can be degraded.  Truth = the same network in fp64.  Error metric = max|a−b| / max|b| per tensor.

Formats: tf32_rn (10-bit mantissa, round to nearest — operands pre-rounded by the producer), tf32_tr (10-bit, truncated
— what tcgen05 kind::tf32 does with raw fp32 operands), bf16 (single pass), bf16x3 (hi/lo split, lo·lo dropped: this
repo's parity mode), fp32.

    python scripts/precision_table.py [B]   →  markdown table on stdout (profiles/r02_precision_table.md)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lxrt_oracle as O  # noqa: E402
from xlxmert_b200 import params as P, synth  # noqa: E402
from xlxmert_b200.config import DEFAULT_DIMS as D  # noqa: E402


def rnd(x, fmt):
    if fmt in ("fp64",):
        return x
    x32 = x.float()
    if fmt == "fp32":
        return x32.double()
    if fmt == "bf16":
        return x32.bfloat16().double()
    if fmt == "bf16x3":                    # handled at the product level
        return x32.double()
    bits = x32.view(torch.int32)
    if fmt == "tf32_tr":
        return (bits & ~0x1FFF).view(torch.float32).double()
    if fmt == "tf32_rn":
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32).double()
    raise ValueError(fmt)


def prod(a, b, fmt):
    """a @ b in fp64 from operands in format `fmt` (fp32 accumulate error is not modelled: it is ≈ 1e-6)."""
    if fmt == "bf16x3":
        a32, b32 = a.float(), b.float()
        ah, bh = a32.bfloat16().float(), b32.bfloat16().float()
        al, bl = (a32 - ah).bfloat16().double(), (b32 - bh).bfloat16().double()
        ah, bh = ah.double(), bh.double()
        return ah @ bh + al @ bh + ah @ bl
    return rnd(a, fmt) @ rnd(b, fmt)


class QLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, fmts):
        ctx.save_for_backward(x, w)
        ctx.fmts = fmts
        return prod(x, w.t(), fmts[0])

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        f = ctx.fmts
        dx = prod(dy, w, f[1])
        dw = prod(dy.reshape(-1, dy.shape[-1]).t(), x.reshape(-1, x.shape[-1]), f[2])
        return dx, dw, None


def run(fmts_by_stream, B, seed=0):
    """fmts_by_stream: dict 'lang'/'vis' → (fwd, dgrad, wgrad) formats.  Returns outputs and gradients (fp64)."""
    d = D
    sd = P.init_state_dict(P.model_param_specs(d), seed=seed, randomize_ln_bias=True)
    batch = synth.make_batch(d, B, 20, 64, seed=seed)
    feats = synth.visual_feats_from(synth.centroid_table(d), batch["cluster_ids"]).double()
    emb = O.embeddings(O.sub(sd, "embeddings"), batch["input_ids"]).double().requires_grad_(True)
    mask = O.extended_mask(batch["attention_mask"], torch.float64)
    sdo = {k: v.double().requires_grad_(True) for k, v in O.sub(sd, "encoder").items()}
    orig = O.linear

    def qlinear(x, w, b):
        # language-side tensors have 20 tokens, vision-side 64 (the joint cross-attention projections see both)
        stream = "lang" if x.shape[-2] == 20 else "vis"
        y = QLinear.apply(x, w, fmts_by_stream[stream])
        return y if b is None else y + b
    O.linear = qlinear
    try:
        ls, vs = O.encoder(sdo, emb, mask, feats, batch["visual_pos"].double(), None, heads=d.heads, n_l=d.l_layers,
                           n_r=d.r_layers, n_x=d.x_layers)
        g = torch.Generator().manual_seed(seed + 77)
        pl, pv = torch.randn(ls[-1].shape, generator=g).double(), torch.randn(vs[-1].shape, generator=g).double()
        ((ls[-1] * pl).sum() + (vs[-1] * pv).sum()).backward()
    finally:
        O.linear = orig
    grads = {k: v.grad for k, v in sdo.items()}
    return dict(lang=ls[-1].detach(), vis=vs[-1].detach(), d_emb=emb.grad, grads=grads)


def err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count())
    truth = run({"lang": ("fp64",) * 3, "vis": ("fp64",) * 3}, B)

    def row(name, fl, fv):
        r = run({"lang": fl, "vis": fv}, B)
        ge = [(err(r["grads"][k], truth["grads"][k]), k) for k in truth["grads"]
              if not k.endswith("key.bias") and float(truth["grads"][k].abs().max()) > 1e-12]
        worst = max(ge)
        wg = max(e for e, k in ge if k.endswith("weight") and "LayerNorm" not in k and "layer_norm" not in k)
        print(f"| {name} | {err(r['lang'], truth['lang']):.1e} | {err(r['vis'], truth['vis']):.1e} | "
              f"{err(r['d_emb'], truth['d_emb']):.1e} | {wg:.1e} | {worst[0]:.1e} ({worst[1]}) |", flush=True)

    print(f"Encoder fwd+bwd, B={B}, L=20, V=64, seeded synthetic inputs; truth = fp64; error = max|a-b|/max|b| per tensor\n")
    print("| operand format of the Linear products (fwd / dgrad / wgrad) | lang out | vis out | d(embedding) | worst Linear-weight grad | worst parameter grad |")
    print("|---|---|---|---|---|---|")
    f32 = ("fp32",) * 3
    for fmt in ("fp32", "bf16x3", "tf32_rn", "tf32_tr", "bf16"):
        row(f"all products {fmt}", (fmt,) * 3, (fmt,) * 3)
    for fmt in ("tf32_rn", "tf32_tr", "bf16"):
        row(f"only WGRAD {fmt} (fwd/dgrad bf16x3)", ("bf16x3", "bf16x3", fmt), ("bf16x3", "bf16x3", fmt))
        row(f"only DGRAD {fmt}", ("bf16x3", fmt, "bf16x3"), ("bf16x3", fmt, "bf16x3"))
        row(f"only FWD {fmt}", (fmt, "bf16x3", "bf16x3"), (fmt, "bf16x3", "bf16x3"))
    row("language stream tf32_rn everywhere, vision bf16x3", ("tf32_rn",) * 3, ("bf16x3",) * 3)
    row("vision stream tf32_rn everywhere, language bf16x3", ("bf16x3",) * 3, ("tf32_rn",) * 3)


if __name__ == "__main__":
    main()
