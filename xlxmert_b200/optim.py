"""Fused clip + AdamW step on the sm_100a library — drop-in for the optimiser half of the reference's training step
(``x-lxmert/src/pretrain/lxmert_pretrain.py:343-364``): ``clip_grad_norm_(model.parameters(), 1.0)`` followed by
``transformers.optimization.AdamW.step()`` (HF 4.1.1 semantics: eps 1e-6, ``correct_bias=True``, weight decay applied
after the Adam update with the plain learning rate), two kernel launches for the whole model.

``B200AdamW`` is a ``torch.optim.Optimizer`` (param groups, ``state_dict``, LR schedulers all work);
``step(max_grad_norm=…)`` fuses the clipping.  Parameters whose ``.grad`` is ``None`` are skipped, like in torch.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from . import _lib


def lxmert_param_groups(model: torch.nn.Module, weight_decay: float):
    """The reference's two groups (lxmert_pretrain.py:122-135): names containing ``bias`` or ``LayerNorm.weight`` get no
    decay (note: by substring — ``visn_layer_norm.weight`` *is* decayed)."""
    no_decay = ["bias", "LayerNorm.weight"]
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    return [
        {"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
        {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
                 correct_bias: bool = True):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self._scratch = None
        self._sqnorm = None

    @staticmethod
    def _arr(ptrs: List[int]):
        return (C.c_void_p * len(ptrs))(*ptrs)

    @torch.no_grad()
    def grad_sqnorm(self, params: Optional[Iterable[torch.Tensor]] = None) -> torch.Tensor:
        """Σ‖g‖² over every parameter that has a gradient (0-d device tensor): ``clip_grad_norm_``'s total norm²."""
        lib = _lib.load()
        ps = [p for g in self.param_groups for p in g["params"]] if params is None else list(params)
        gs = [p.grad for p in ps if p.grad is not None]
        if not gs:
            raise RuntimeError("no gradients")
        for g in gs:
            if not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous():
                raise TypeError("B200AdamW needs contiguous fp32 CUDA gradients (no CPU fallback)")
        dev = gs[0].device
        elems = (C.c_int64 * len(gs))(*[g.numel() for g in gs])
        need = lib.xlx_optim_scratch_floats(elems, len(gs))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, device=dev, dtype=torch.float32)
            self._sqnorm = torch.empty((), device=dev, dtype=torch.float32)
        rc = lib.xlx_grad_sqnorm(self._arr([g.data_ptr() for g in gs]), elems, len(gs), self._scratch.data_ptr(),
                                 self._sqnorm.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check("xlx_grad_sqnorm", rc)
        return self._sqnorm

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: float = 0.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        from .encoder import finish_pending_gradient_sync
        finish_pending_gradient_sync()       # gradient all-reduces a data-parallel backward may have left in flight
        sq = self.grad_sqnorm() if max_grad_norm > 0 else None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            steps = set()
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                steps.add(st["step"])
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise TypeError("B200AdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
            b1, b2 = group["betas"]
            # tensors that joined later (their first gradient) have their own step count: one launch per distinct count
            for s in sorted(steps):
                sel = [p for p in ps if self.state[p]["step"] == s]
                n = len(sel)
                elems = (C.c_int64 * n)(*[p.numel() for p in sel])
                wd = (C.c_float * n)(*([float(group["weight_decay"])] * n))
                rc = lib.xlx_adamw_step(self._arr([p.data_ptr() for p in sel]), self._arr([p.grad.data_ptr() for p in sel]),
                                        self._arr([self.state[p]["exp_avg"].data_ptr() for p in sel]),
                                        self._arr([self.state[p]["exp_avg_sq"].data_ptr() for p in sel]), elems, wd, n,
                                        float(group["lr"]), float(b1), float(b2), float(group["eps"]), int(s),
                                        int(bool(group["correct_bias"])), None if sq is None else sq.data_ptr(),
                                        float(max_grad_norm), torch.cuda.current_stream().cuda_stream)
                _lib.check("xlx_adamw_step", rc)
            # The kernels wrote through raw pointers, behind autograd's back.  Bump the version counters so that every
            # cache keyed on (data_ptr, _version) — the split-bf16 weight copies of the encoder, heads and generator —
            # is rebuilt before the next forward.  (`p.data = p.data` does NOT change `_version`.)
            torch.autograd.graph.increment_version(ps)
        return loss
