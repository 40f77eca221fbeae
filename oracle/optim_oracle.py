"""ORACLE — test infrastructure, not product code.

CPU restatement of the optimiser half of the reference's training step
(``/root/reference/x-lxmert/src/pretrain/lxmert_pretrain.py:343-364``):

* ``torch.nn.utils.clip_grad_norm_(parameters, max_norm)`` — ``total_norm = ‖(‖g_i‖₂)_i‖₂``,
  ``coef = max_norm / (total_norm + 1e-6)``, gradients scaled by ``min(coef, 1)``;
* ``transformers.optimization.AdamW.step`` as published in transformers **4.1.1** (the version pinned in
  ``/root/reference/requirements.txt:11``; the class was removed from the 5.5.0 installed here, so its source is not
  available in this container and this restatement follows the published algorithm — parity for this row is
  therefore "unpinned" by reference-run goldens; ``tests/test_optim_oracle.py`` cross-checks it on CPU against
  ``torch.nn.utils.clip_grad_norm_`` and against ``torch.optim.AdamW`` where the two algorithms coincide (eps = 0, no
  weight decay), plus a hand-computed step for the decay placement).

      exp_avg    ← β₁·exp_avg + (1 − β₁)·g
      exp_avg_sq ← β₂·exp_avg_sq + (1 − β₂)·g²
      denom      ← sqrt(exp_avg_sq) + eps                       (eps 1e-6 by default)
      step_size  ← lr·sqrt(1 − β₂ᵗ)/(1 − β₁ᵗ)   if correct_bias else lr
      p          ← p − step_size·exp_avg/denom
      p          ← p − lr·weight_decay·p                        (after the Adam update, plain lr)
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch


def clip_coef(grads: List[torch.Tensor], max_norm: float) -> float:
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    c = max_norm / (total + 1e-6)
    return min(c, 1.0)


def adamw_step(p: torch.Tensor, g: torch.Tensor, state: Dict, lr: float, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-6, weight_decay: float = 0.0, correct_bias: bool = True) -> None:
    """In-place HF-AdamW update of ``p`` (fp32 arithmetic like the reference)."""
    if not state:
        state["step"] = 0
        state["exp_avg"] = torch.zeros_like(p)
        state["exp_avg_sq"] = torch.zeros_like(p)
    state["step"] += 1
    m, v = state["exp_avg"], state["exp_avg_sq"]
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = lr * math.sqrt(1.0 - beta2 ** state["step"]) / (1.0 - beta1 ** state["step"])
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
