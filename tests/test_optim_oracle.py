"""CPU checks of the optimiser oracle (oracle/optim_oracle.py).  The HF AdamW class it restates no longer exists in the
installed transformers, so the restatement is cross-checked where independent implementations coincide with it:
``torch.nn.utils.clip_grad_norm_`` for the clip coefficient, and ``torch.optim.AdamW`` in the regime where the two
algorithms are algebraically identical (eps = 0, no weight decay)."""
import torch

from oracle import optim_oracle as OO


def test_clip_coefficient_matches_torch():
    g = torch.Generator().manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((7, 5), (11,), (3, 4, 2))]
    for scale in (0.01, 10.0):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g) * scale
        before = [p.grad.clone() for p in params]
        coef = OO.clip_coef(before, 1.0)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        for p, b in zip(params, before):
            assert torch.allclose(p.grad, b * coef, rtol=1e-6, atol=1e-9)
        assert (coef < 1.0) == (scale > 1.0)


def test_adamw_restatement_equals_torch_adamw_where_the_algorithms_coincide():
    g = torch.Generator().manual_seed(1)
    p_ref = torch.nn.Parameter(torch.randn(64, 33, generator=g))
    p_orc = p_ref.detach().clone()
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), eps=0.0, weight_decay=0.0)
    state = {}
    for _ in range(6):
        grad = torch.randn(p_ref.shape, generator=g) + 0.1          # non-zero everywhere: eps = 0 is safe
        p_ref.grad = grad.clone()
        opt.step()
        OO.adamw_step(p_orc, grad, state, lr=1e-3, eps=0.0, weight_decay=0.0, correct_bias=True)
    assert torch.allclose(p_orc, p_ref.detach(), rtol=2e-6, atol=1e-7)
    st = opt.state[p_ref]
    # torch updates exp_avg with lerp_ (m + (g − m)·(1 − β₁)): same value, different rounding
    assert torch.allclose(state["exp_avg"], st["exp_avg"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(state["exp_avg_sq"], st["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_weight_decay_is_applied_after_the_update_with_the_plain_learning_rate():
    """The one place HF's AdamW differs structurally from torch's: p ← p_adam − lr·wd·p_adam (optimization.py, 4.1.1)."""
    p = torch.ones(4)
    g = torch.full((4,), 0.5)
    state = {}
    OO.adamw_step(p, g, state, lr=0.1, eps=1e-6, weight_decay=0.01, correct_bias=True)
    # first step: m = 0.05, v = 0.00025, step_size = lr·sqrt(1−β2)/(1−β1), update = step_size·m/(sqrt(v)+eps)
    m, v = 0.05, 0.00025
    step_size = 0.1 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    p1 = 1.0 - step_size * m / (v ** 0.5 + 1e-6)
    want = p1 - 0.1 * 0.01 * p1
    assert torch.allclose(p, torch.full((4,), want), rtol=1e-6)
