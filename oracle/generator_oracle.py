"""ORACLE — test infrastructure, not product code.

CPU restatement of the grid-feature GAN generator forward,
``/root/reference/image_generator/src/layers.py:9-260`` (``SPADE`` :9-47, ``NoiseInjection`` :50-62,
``GeneratorResidualBlock`` :65-113, ``ToRGB`` :116-132, ``Generator`` :135-260), in *eval* mode:
spectral-normalised convolutions use the frozen effective weight ``weight_orig / σ`` with
``σ = uᵀ·W_mat·v`` (torch legacy ``spectral_norm`` hook, SURVEY App. A "Generator").

Resampling is restated explicitly (``bilinear_resize`` below) so that the CUDA kernels have an exact
index/weight definition to match: ``align_corners=False`` source coordinate
``src = max((dst + 0.5)·in/out − 0.5, 0)``, ``i0 = floor(src)``, ``i1 = min(i0 + 1, in − 1)``,
``w1 = src − i0`` — ATen's ``area_pixel_compute_source_index`` / ``upsample_bilinear2d``.

Pinned by ``tests/golden/generator_*.npz`` (outputs of the reference class itself, produced by
``oracle/make_golden.py``) — see ``tests/test_oracle_golden.py``.  Only tests, ``smoke()`` and the
bench's CPU leg may import this file.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _axis_weights(n_in: int, n_out: int, dtype, device):
    scale = n_in / n_out
    dst = torch.arange(n_out, dtype=dtype, device=device)
    src = ((dst + 0.5) * scale - 0.5).clamp(min=0)
    i0 = src.floor().to(torch.int64).clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    w1 = src - i0.to(dtype)
    return i0, i1, 1.0 - w1, w1


def bilinear_resize(x: Tensor, size: int) -> Tensor:
    """``F.interpolate(x, size=(size,size), mode='bilinear', align_corners=False)`` restated."""
    _, _, H, W = x.shape
    y0, y1, wy0, wy1 = _axis_weights(H, size, x.dtype, x.device)
    x0, x1, wx0, wx1 = _axis_weights(W, size, x.dtype, x.device)
    rows = x[:, :, y0, :] * wy0[None, None, :, None] + x[:, :, y1, :] * wy1[None, None, :, None]
    return rows[:, :, :, x0] * wx0 + rows[:, :, :, x1] * wx1


def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """``InstanceNorm2d(affine=False)``: biased variance over H×W per (sample, channel)."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=(2, 3), keepdim=True)
    return xc * torch.rsqrt(var + eps)


def sn_weight(sd: SD, prefix: str) -> Tensor:
    """Effective eval-mode weight of a spectral-normalised conv: ``weight_orig / (uᵀ W v)``."""
    w = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    sigma = torch.dot(u, w.reshape(w.shape[0], -1) @ v)
    return w / sigma


def spade(sd: SD, p: str, x: Tensor, y: Tensor) -> Tensor:
    """``SPADE.forward`` (layers.py:33-47)."""
    normalized = instance_norm(x)
    ys = bilinear_resize(y, x.shape[2])
    actv = F.relu(F.conv2d(ys, sd[p + ".shared.0.weight"], sd[p + ".shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[p + ".gamma.weight"], sd[p + ".gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[p + ".beta.weight"], sd[p + ".beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def resblock(sd: SD, p: str, x: Tensor, y: Tensor) -> Tensor:
    """``GeneratorResidualBlock.forward`` (layers.py:93-113), noise off."""
    h = spade(sd, p + ".cbn1", x, y)
    h = F.leaky_relu(h, 0.2)
    h = bilinear_resize(h, 2 * h.shape[2])
    h = F.conv2d(h, sn_weight(sd, p + ".conv1"), sd[p + ".conv1.bias"], padding=1)
    h = spade(sd, p + ".cbn2", h, y)
    h = F.leaky_relu(h, 0.2)
    h = F.conv2d(h, sn_weight(sd, p + ".conv2"), sd[p + ".conv2.bias"], padding=1)
    res = bilinear_resize(x, 2 * x.shape[2])
    res = F.conv2d(res, sn_weight(sd, p + ".res_branch.1"), sd[p + ".res_branch.1.bias"])
    return h + res


def generator(sd: SD, emb: Tensor, target_size: int = 256, return_intermediates: bool = False):
    """``Generator.forward(emb, train=False)`` (layers.py:223-253).

    ``emb``: ``[B, 2048, 8, 8]`` or ``[B, 8, 8, 2048]`` (permuted as at :231-233).
    Returns the image ``[B, 3, T, T]`` in (−1, 1); with ``return_intermediates`` also the per-block
    ``h`` maps and the pre-tanh accumulator."""
    if emb.shape[1:] == (8, 8, sd["bottleneck_emb.0.weight"].shape[1]):
        emb = emb.permute(0, 3, 1, 2)
    e = torch.tanh(F.conv2d(emb, sd["bottleneck_emb.0.weight"], sd["bottleneck_emb.0.bias"]))
    h = F.conv2d(e, sn_weight(sd, "learned_init_conv.0"), sd["learned_init_conv.0.bias"], padding=1,
                 groups=4)
    y = F.conv2d(e, sn_weight(sd, "style_init_conv.0"), sd["style_init_conv.0.bias"], padding=1,
                 groups=4)
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("resblocks.")})
    out = torch.zeros(emb.shape[0], 3, target_size, target_size, dtype=emb.dtype)
    hs = []
    for i in range(n_blocks):
        h = resblock(sd, f"resblocks.{i}", h, y)
        hs.append(h)
        rgb = F.conv2d(h, sd[f"to_RGB_blocks.{i}.conv.weight"], sd[f"to_RGB_blocks.{i}.conv.bias"],
                       padding=1)
        if i + 1 < n_blocks:
            rgb = bilinear_resize(rgb, target_size)
        out = out + rgb
    img = torch.tanh(out)
    if return_intermediates:
        return img, dict(h=hs, pre_tanh=out, style=y, bottleneck=e)
    return img
