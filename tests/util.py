"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def checksum(sd) -> float:
    return float(sum((v.double().abs().sum() + v.double().sum() * 0.5) for v in sd.values()
                     if v.is_floating_point()))


def rel_err(a, b):
    """Tensor-normalised max error: max|a-b| / max|b| (SURVEY §7.2-1 recommended metric)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def probes(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


def assert_argmax_stratified(pred_id, ref_id, ref_top2, ref_margin, tau, what=""):
    """Index parity bar for a cluster arg-max (north_star: bit-exact): every row whose REFERENCE top-2 logit margin
    exceeds ``tau`` must match the reference index exactly; a row below it (the reference's own fp32 arithmetic does not
    resolve it: summation-order noise is ≈ 5e-6) must pick one of the reference's top-2.  Prints the
    flip count so that the log of a green run still shows how many near-tie rows went the other way."""
    pred = torch.as_tensor(pred_id).cpu().long().reshape(-1)
    ref = torch.as_tensor(ref_id).cpu().long().reshape(-1)
    top2 = torch.as_tensor(ref_top2).cpu().long().reshape(-1, 2)
    margin = torch.as_tensor(ref_margin).cpu().double().reshape(-1)
    clear = margin > tau
    flips = pred != ref
    print(f"[argmax {what}] rows {pred.numel()}, clear(margin>{tau:g}) {int(clear.sum())}, "
          f"flips {int(flips.sum())} (all inside the near-tie band: {bool((~clear[flips]).all()) if flips.any() else True})")
    assert torch.equal(pred[clear], ref[clear]), \
        f"{what}: {int((flips & clear).sum())} arg-max flips on rows with margin > {tau:g}"
    near = ~clear
    member = (pred[near] == top2[near, 0]) | (pred[near] == top2[near, 1])
    assert bool(member.all()), f"{what}: a near-tie row picked an index outside the reference's top-2"
    return int(flips.sum())
