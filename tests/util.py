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
