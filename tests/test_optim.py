"""GPU parity of the fused clip + AdamW step against the CPU restatement of the reference's optimiser step
(oracle/optim_oracle.py).  Tolerance: 2e-6 relative on parameters and moments after several steps (fp32 elementwise
arithmetic; the only differences are FMA contraction and the order of the norm's summation)."""
import math

import pytest
import torch

from oracle import optim_oracle as OO

pytestmark = pytest.mark.gpu


def _setup(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(s, generator=g) for s in shapes]
    return ps, g


def test_fused_adamw_with_clipping_matches_oracle_over_steps():
    from xlxmert_b200.optim import B200AdamW
    shapes = [(768, 768), (3072,), (5, 7), (1,), (4099,), (30522, 8)]
    ps_cpu, gen = _setup(0, shapes)
    ps_gpu = [torch.nn.Parameter(p.clone().cuda()) for p in ps_cpu]
    wd = [0.01, 0.0, 0.01, 0.0, 0.01, 0.01]
    opt = B200AdamW([{"params": [ps_gpu[i] for i in range(6) if wd[i] > 0], "weight_decay": 0.01},
                     {"params": [ps_gpu[i] for i in range(6) if wd[i] == 0], "weight_decay": 0.0}], lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    states = [dict() for _ in shapes]
    for step in range(4):
        scale = 10.0 if step == 0 else 0.01          # step 0 clips, later steps do not
        grads = [scale * torch.randn(s, generator=gen) for s in shapes]
        if step == 2:
            grads[4] = None                          # a parameter without gradient this step is skipped
        live = [g for g in grads if g is not None]
        coef = OO.clip_coef(live, 1.0)
        lr = 1e-3 / (1 + step)
        for p, g, st, w in zip(ps_cpu, grads, states, wd):
            if g is not None:
                OO.adamw_step(p, g * coef, st, lr, weight_decay=w)
        for p, g in zip(ps_gpu, grads):
            p.grad = None if g is None else g.clone().cuda()
        opt.step(max_grad_norm=1.0)
        sched.step()
        total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in live))
        assert abs(math.sqrt(float(opt.grad_sqnorm())) - total) < 1e-5 * total
    for p, q, st in zip(ps_gpu, ps_cpu, states):
        assert float((p.detach().cpu() - q).abs().max()) < 2e-6 * float(q.abs().max())
        assert float((opt.state[p]["exp_avg"].cpu() - st["exp_avg"]).abs().max()) < 2e-6 * float(st["exp_avg"].abs().max())
        assert float((opt.state[p]["exp_avg_sq"].cpu() - st["exp_avg_sq"]).abs().max()) < 2e-6 * float(st["exp_avg_sq"].abs().max())
        assert opt.state[p]["step"] == st["step"]


def test_fused_adamw_matches_plain_adam_arithmetic_without_decay():
    """Independent cross-check (not through the oracle): with no decay and no clipping one step equals
    p − lr·sqrt(1−β₂)/(1−β₁)·(1−β₁)g / (sqrt((1−β₂)g²) + eps)."""
    from xlxmert_b200.optim import B200AdamW
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g)
    gr = torch.randn(1000, generator=g)
    p = torch.nn.Parameter(p0.clone().cuda())
    p.grad = gr.clone().cuda()
    B200AdamW([p], lr=0.01, eps=1e-6).step()
    m, v = 0.1 * gr.double(), 0.001 * gr.double() ** 2
    want = p0.double() - 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9) * m / (v.sqrt() + 1e-6)
    assert float((p.detach().cpu().double() - want).abs().max()) < 1e-6


def test_fused_adamw_rejects_cpu_parameters():
    from xlxmert_b200.optim import B200AdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(TypeError):
        B200AdamW([p]).step()


def test_optimizer_step_invalidates_prepared_weight_caches():
    """the fused step writes through raw pointers; the split-bf16 weight copies of the native modules
    are keyed on (data_ptr, _version), so the step must bump the version counters.  forward → step → forward must
    change the output and equal a model rebuilt from the updated weights."""
    from test_pretrain_parity import build_model
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as d
    from xlxmert_b200.optim import B200AdamW

    model, table = build_model(5)
    model.train()
    batch = {k: v.cuda() for k, v in synth.make_batch(d, 2, 12, 64, seed=1).items()}
    kw = dict(input_ids=batch["input_ids"], visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
              cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"], task="vis_mask",
              label_dict={"obj_labels": batch["obj_labels"]})
    versions = {n: p._version for n, p in model.named_parameters()}
    opt = B200AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    l0 = model(**kw)["total_loss"]
    l0.backward()
    stepped = [n for n, p in model.named_parameters() if p.grad is not None]
    assert stepped
    opt.step(max_grad_norm=1.0)
    for n, p in model.named_parameters():
        if n in stepped:
            assert p._version > versions[n], n
    for p in model.parameters():
        p.grad = None
    with torch.no_grad():
        l1 = model(**kw)["total_loss"]
    assert abs(float(l1) - float(l0)) > 1e-4 * abs(float(l0)), "the forward after the step still sees the old weights"
    fresh, _ = build_model(5)
    fresh.load_state_dict({k: v.detach().clone() for k, v in model.state_dict().items()}, strict=True)
    fresh.train()
    with torch.no_grad():
        l2 = fresh(**kw)["total_loss"]
    assert abs(float(l2) - float(l1)) <= 1e-6 * abs(float(l1)), (float(l1), float(l2))
