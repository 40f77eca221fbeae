"""GPU parity of B200LxmertModel (embeddings → encoder → pooler, all through the C ABI) against the goldens the
reference's own HF ``LxmertModel`` produced (oracle/make_golden.py: forward outputs, probe loss and the
gradient of every parameter).  Tolerances as in test_encoder_parity.py (1e-4 outputs, 1e-3 gradients)."""
import numpy as np
import pytest
import torch

from xlxmert_b200 import params as P
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D

from util import load_golden, probes, rel_err

pytestmark = pytest.mark.gpu


def _model(sd, **kw):
    from xlxmert_b200.lxmert import B200LxmertModel
    m = B200LxmertModel(D, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.cuda()


@pytest.mark.parametrize("name", ["model_b2_l20_v64", "model_b3_l13_v36"])
def test_model_forward_backward_matches_reference_golden(name):
    g = load_golden(name)
    B, L, V, wseed, bseed = (int(x) for x in g["meta"])
    sd = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    batch = synth.make_batch(D, B, L, V, seed=bseed)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"]).cuda().requires_grad_(True)
    m = _model(sd).train()
    out = m(input_ids=batch["input_ids"].cuda(), visual_feats=feats, visual_pos=batch["visual_pos"].cuda(),
            attention_mask=batch["attention_mask"].cuda())
    lang, vis, pooled = out[0], out[1], out[2]
    assert out.language_output is lang and out.pooled_output is pooled
    assert rel_err(lang.detach().cpu(), g["lang"]) < 1e-4
    assert rel_err(vis.detach().cpu(), g["vis"]) < 1e-4
    assert rel_err(pooled.detach().cpu(), g["pooled"]) < 1e-4
    pl, pv, pp = probes([lang.shape, vis.shape, pooled.shape], seed=bseed + 77)
    loss = (lang * pl.cuda()).sum() + (vis * pv.cuda()).sum() + (pooled * pp.cuda()).sum()
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    assert rel_err(feats.grad.cpu()[:, ::8, ::64], g["dfeats_sub"]) < 1e-3
    named = dict(m.named_parameters())
    checked = 0
    for k, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        k = str(k)
        p = named[k]
        assert p.grad is not None, k
        gn = float(p.grad.double().norm())
        if norm < 1e-5:                       # mathematically zero (key bias; padding rows)
            assert gn < 1e-4, (k, gn)
            continue
        assert abs(gn - norm) < 2e-3 * norm, (k, gn, norm)
        n = min(8, p.numel())
        got = p.grad.flatten()[:n].double().cpu().numpy()
        scale = max(float(np.abs(head[:n]).max()), norm / np.sqrt(p.numel()))
        assert float(np.abs(got - head[:n]).max()) < 2e-3 * scale, (k, got, head[:n])
        checked += 1
    assert checked > 300
    # padding_idx rows receive no gradient (HF:184-186)
    for t in ("word", "position", "token_type"):
        assert float(named[f"embeddings.{t}_embeddings.weight"].grad[0].abs().max()) == 0.0


def test_model_hidden_states_and_eval_mode():
    g = load_golden("model_b2_l20_v64")
    B, L, V, wseed, bseed = (int(x) for x in g["meta"])
    sd = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    batch = synth.make_batch(D, B, L, V, seed=bseed)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"]).cuda()
    m = _model(sd).eval()
    with torch.no_grad():
        out = m(input_ids=batch["input_ids"].cuda(), visual_feats=feats, visual_pos=batch["visual_pos"].cuda(),
                attention_mask=batch["attention_mask"].cuda(), output_hidden_states=True)
    assert len(out.language_hidden_states) == 14 and len(out.vision_hidden_states) == 10
    assert rel_err(out.language_hidden_states[3].cpu()[:, ::4, ::8], g["lang_h3"]) < 1e-4
    assert rel_err(out[2].cpu(), g["pooled"]) < 1e-4
    with pytest.raises(ValueError):
        m(input_ids=batch["input_ids"].cuda())
