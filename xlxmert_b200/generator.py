"""Drop-in for ``Generator`` (``image_generator/src/layers.py:135-260``) — the module the sampler calls as
``self.G(code.permute(0,2,1).view(B,2048,8,8))`` (``x-lxmert/src/tasks/imggen_model.py:254``) and that
``tasks/sample_images.py:53-69`` builds and loads ``G_60.pth`` into.

Same constructor keywords, same state-dict keys (``*.weight_orig`` / ``weight_u`` / ``weight_v`` for the spectrally
normalised convs, plain ``weight`` for the SPADE / ToRGB / bottleneck convs) and ``forward(emb, train=True)``.
Only the canonical architecture is supported (``base_dim=32, emb_dim=2048, codebook_dim=256, target_size=256,
init_H=init_W=8, norm_type='spade_in', SN=True``); anything else raises.  Spectral norm has eval semantics
(``weight_orig / σ`` with the stored ``u``/``v``): training the generator is outside the reference's published code
(its trainer is missing).  There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
from torch import nn

from . import _lib

N_BLOCKS = 5


class _Conv(nn.Module):
    """Parameter holder with ``nn.Conv2d``'s names; ``sn`` adds the legacy spectral-norm layout."""

    def __init__(self, cin, cout, k, groups=1, sn=False):
        super().__init__()
        w = torch.empty(cout, cin // groups, k, k)
        nn.init.orthogonal_(w)
        if sn:
            self.weight_orig = nn.Parameter(w)
            self.register_buffer("weight_u", nn.functional.normalize(torch.randn(cout), dim=0))
            self.register_buffer("weight_v", nn.functional.normalize(torch.randn(w[0].numel()), dim=0))
        else:
            self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(cout))
        self.sn = sn

    def tensors(self):
        if self.sn:
            return [self.weight_orig, self.bias, self.weight_u, self.weight_v]
        return [self.weight, self.bias]


class _Spade(nn.Module):            # layers.py:9-31
    def __init__(self, x_dim, y_dim):
        super().__init__()
        self.shared = nn.Sequential(_Conv(y_dim, 128, 3))
        self.gamma = _Conv(128, x_dim, 3)
        self.beta = _Conv(128, x_dim, 3)

    def tensors(self):
        return self.shared[0].tensors() + self.gamma.tensors() + self.beta.tensors()


class _Noise(nn.Module):            # layers.py:50-54
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class _ResBlock(nn.Module):         # layers.py:65-91
    def __init__(self, c):
        super().__init__()
        self.cbn1 = _Spade(c, c)
        self.conv1 = _Conv(c, c, 3, sn=True)
        self.noise1 = _Noise()
        self.cbn2 = _Spade(c, c)
        self.conv2 = _Conv(c, c, 3, sn=True)
        self.noise2 = _Noise()
        self.res_branch = nn.Sequential(nn.Identity(), _Conv(c, c, 1, sn=True))

    def tensors(self):
        return (self.cbn1.tensors() + self.cbn2.tensors() + self.conv1.tensors() + self.conv2.tensors()
                + self.res_branch[1].tensors() + [self.noise1.weight, self.noise2.weight])


class _ToRGB(nn.Module):            # layers.py:116-124
    def __init__(self, c):
        super().__init__()
        self.conv = _Conv(c, 3, 3)


class B200Generator(nn.Module):
    def __init__(self, emb_dim=2048, mod_dim=128, base_dim=32, n_channel=3, target_size=256, extra_layers=0,
                 init_H=8, init_W=8, norm_type='spade_in', SN=True, codebook_dim=256, passes: int = 3):
        super().__init__()
        if (emb_dim, base_dim, n_channel, target_size, extra_layers, init_H, init_W, norm_type, bool(SN),
                codebook_dim) != (2048, 32, 3, 256, 0, 8, 8, 'spade_in', True, 256):
            raise NotImplementedError("B200Generator implements the canonical X-LXMERT generator only "
                                      "(base_dim 32, emb_dim 2048, codebook_dim 256, 8x8 -> 256x256, spade_in, SN)")
        self.init_H, self.init_W, self.target_size, self.emb_dim = init_H, init_W, target_size, emb_dim
        self.norm_type, self.SN = norm_type, SN
        self.passes = passes
        self.bottleneck_emb = nn.Sequential(_Conv(emb_dim, codebook_dim, 1))
        self.learned_init_conv = nn.Sequential(_Conv(codebook_dim, base_dim, 3, groups=4, sn=True))
        self.style_init_conv = nn.Sequential(_Conv(codebook_dim, base_dim, 3, groups=4, sn=True))
        self.resblocks = nn.ModuleList([_ResBlock(base_dim) for _ in range(N_BLOCKS)])
        self.to_RGB_blocks = nn.ModuleList([_ToRGB(base_dim) for _ in range(N_BLOCKS)])
        self._prep = None
        self._prep_key = None
        self._parr = None

    def _tensors(self) -> List[torch.Tensor]:
        t = self.bottleneck_emb[0].tensors() + self.learned_init_conv[0].tensors() + self.style_init_conv[0].tensors()
        for rb in self.resblocks:
            t += rb.tensors()
        for rgb in self.to_RGB_blocks:
            t += rgb.conv.tensors()
        return t

    def _prepared(self):
        lib = _lib.load()
        ts = self._tensors()
        assert len(ts) == lib.xlx_generator_num_params()
        key = tuple((t.data_ptr(), t._version) for t in ts)
        dev = ts[0].device
        if self._prep is None or self._prep.device != dev:
            self._prep = torch.empty(lib.xlx_generator_prep_bytes(), dtype=torch.uint8, device=dev)
            self._prep_key = None
        if key != self._prep_key:
            for t in ts:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise TypeError("generator parameters must be contiguous fp32")
            self._parr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            rc = lib.xlx_generator_prepare(self._parr, self._prep.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check("xlx_generator_prepare", rc)
            self._prep_key = key
        return self._prep, self._parr

    def _noise_active(self) -> bool:
        """Whether any NoiseInjection weight is non-zero (layers.py:56-62: they are initialised to 0, so a random-init G
        is deterministic).  Reading ten device scalars costs ten host syncs, so the answer is cached per weight version
        — which also keeps the forward capturable in a CUDA graph."""
        ws = [w for rb in self.resblocks for w in (rb.noise1.weight, rb.noise2.weight)]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        cached = self.__dict__.get("_noise_key")
        if cached is None or cached[0] != key:
            cached = (key, bool(torch.cat([w.detach().reshape(-1) for w in ws]).ne(0).any().item()))
            self.__dict__["_noise_key"] = cached
        return cached[1]

    @torch.no_grad()
    def forward(self, emb, train=True, return_intermediates: bool = False, noise=None):
        """``emb``: ``[B, 2048, 8, 8]`` (any strides) or ``[B, 8, 8, 2048]`` → ``[B, 3, 256, 256]`` in (−1, 1).
        ``train`` only gates noise injection (layers.py:56-62, 246), as in the reference.  ``noise`` (optional, with
        ``train=True``): the ten standard-normal maps to inject instead of fresh draws — ``[B, R, R]`` (or
        ``[B, 1, R, R]``) for noise1 / noise2 of each block in order, R = 8, 16, 16, 32, …, 128, 256."""
        lib = _lib.load()
        if not emb.is_cuda:
            raise RuntimeError("B200Generator runs on CUDA (sm_100a) only; there is no CPU fallback")
        B = emb.shape[0]
        if tuple(emb.shape[1:]) == (self.emb_dim, self.init_H, self.init_W):
            emb = emb.permute(0, 2, 3, 1)                       # → [B, 8, 8, 2048] view
        elif tuple(emb.shape[1:]) != (self.init_H, self.init_W, self.emb_dim):
            raise ValueError(f"unexpected generator input shape {tuple(emb.shape)}")
        emb = emb.contiguous().float()                          # no copy when the caller's memory is cell-major
        dev = emb.device
        prep, parr = self._prepared()
        stream = torch.cuda.current_stream().cuda_stream
        noise_arr, noise_keep = None, []
        if train and noise is not None:
            want = [s for i in range(N_BLOCKS) for s in (8 << i, 16 << i)]
            if len(noise) != len(want):
                raise ValueError(f"noise must hold {len(want)} maps")
            for t, r in zip(noise, want):
                if t.numel() != B * r * r:
                    raise ValueError(f"noise map of {t.numel()} elements where [B={B}, {r}, {r}] is expected")
                noise_keep.append(t.to(device=dev, dtype=torch.float32).reshape(B, r, r).contiguous())
            noise_arr = (C.c_void_p * len(noise_keep))(*[t.data_ptr() for t in noise_keep])
        elif train and self._noise_active():
            for i in range(N_BLOCKS):
                r = 8 << i
                noise_keep += [torch.randn(B, r, r, device=dev), torch.randn(B, 2 * r, 2 * r, device=dev)]
            noise_arr = (C.c_void_p * len(noise_keep))(*[t.data_ptr() for t in noise_keep])
        img = torch.empty(B, 3, 256, 256, device=dev, dtype=torch.float32)
        pre = torch.empty_like(img) if return_intermediates else None
        blocks, blocks_arr = [], None
        if return_intermediates:
            blocks = [torch.empty(B, 16 << i, 16 << i, 32, device=dev) for i in range(N_BLOCKS)]
            blocks_arr = (C.c_void_p * N_BLOCKS)(*[t.data_ptr() for t in blocks])
        nws = lib.xlx_generator_workspace_bytes(B)
        # one grow-only workspace per module: a fresh multi-GB request per call fragments the caching allocator until it
        # falls back to cudaMalloc (tens of ms); stream-ordered reuse is safe for this inference-only module
        ws = self.__dict__.get("_ws")
        if ws is None or ws.numel() < nws or ws.device != dev:
            ws = self.__dict__["_ws"] = torch.empty(nws, dtype=torch.uint8, device=dev)
        rc = lib.xlx_generator_fwd(parr, prep.data_ptr(), B, emb.data_ptr(), noise_arr, img.data_ptr(),
                                   None if pre is None else pre.data_ptr(), blocks_arr, ws.data_ptr(), nws, self.passes,
                                   stream)
        _lib.check("xlx_generator_fwd", rc)
        if return_intermediates:
            return img, pre, [b.permute(0, 3, 1, 2) for b in blocks]     # NCHW views like the reference's h
        return img
