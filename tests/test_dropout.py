"""Training-mode dropout of the native path (VERDICT r1 missing-1; the reference trains under ``model.train()`` with
p = 0.1, lxmert_pretrain.py:271, HF modeling_lxmert.py:189,236,282,344,474).

PyTorch's random stream cannot be reproduced from fused kernels, so equivalence of the MASKS is statistical; what is
checked exactly is the arithmetic around them: the product's counter-based masks are materialised (``xlx_dropout_mask``)
and handed to the CPU oracle, whose forward AND backward must then agree with the CUDA path at the usual tolerances
(1e-4 outputs, 1e-3 gradients) — this pins every site's position in the graph, the 1/(1−p) scaling, the mask replay in
the backward, and the split of the LayerNorm-input gradient into its residual and dense paths."""
import ctypes as C
import math
from dataclasses import replace

import pytest
import torch

from oracle import lxrt_oracle as O
from xlxmert_b200 import params as P, synth
from xlxmert_b200.config import DEFAULT_DIMS

from util import probes, rel_err

pytestmark = pytest.mark.gpu

PH, PA, SEED = 0.1, 0.1, 0x1234ABCD5678
DD = replace(DEFAULT_DIMS, l_layers=2, r_layers=1, x_layers=2, hidden_dropout=PH, attention_dropout=PA)


def _mask(seed, site, p, kind, rows, cols):
    from xlxmert_b200 import _lib
    out = torch.empty(rows, cols, device="cuda", dtype=torch.float32)
    _lib.check("xlx_dropout_mask", _lib.load().xlx_dropout_mask(seed, site, p, kind, rows, cols, out.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream))
    return out


def provider(site, kind, shape):
    if kind == "hidden":
        rows, cols = math.prod(shape[:-1]), shape[-1]
        return _mask(SEED, site, PH, 0, rows, cols).view(shape).cpu()
    rows, cols = math.prod(shape[:-1]), shape[-1]
    return _mask(SEED, site, PA, 1, rows, cols).view(shape).cpu()


@pytest.fixture()
def fixed_seed(monkeypatch):
    from xlxmert_b200 import _lib

    def step_dropout(module, dims):
        if not module.training or (dims.hidden_dropout <= 0 and dims.attention_dropout <= 0):
            return None
        return _lib.XlxDropout(dims.hidden_dropout, dims.attention_dropout, SEED)
    monkeypatch.setattr(_lib, "step_dropout", step_dropout)


def test_mask_statistics_and_determinism():
    for kind, rows, cols, p in ((0, 4096, 768, 0.1), (1, 12 * 64 * 8, 64, 0.1), (1, 12 * 20 * 8, 13, 0.3), (0, 64, 768, 0.5)):
        m = _mask(SEED, 18, p, kind, rows, cols)
        vals = torch.unique(m).tolist()
        assert all(abs(v) < 1e-12 or abs(v - 1 / (1 - p)) < 1e-6 for v in vals), vals
        n = m.numel()
        keep = float((m > 0).float().mean())
        assert abs(keep - (1 - p)) < 5 * math.sqrt(p * (1 - p) / n), (kind, keep)
        assert abs(float(m.mean()) - 1.0) < 6 * math.sqrt(p / (1 - p) / n)                 # unbiased: E[mask] = 1
        assert torch.equal(m, _mask(SEED, 18, p, kind, rows, cols))                         # pure function of its inputs
        assert not torch.equal(m, _mask(SEED, 19, p, kind, rows, cols))                     # sites are independent
        assert not torch.equal(m, _mask(SEED + 1, 18, p, kind, rows, cols))                 # and so are steps
        # no visible correlation between neighbouring elements (Philox groups of 4 / 2)
        a, b = (m.flatten()[:-1] > 0).float(), (m.flatten()[1:] > 0).float()
        assert abs(float((a * b).mean()) - (1 - p) ** 2) < 6 / math.sqrt(n)
    ones = _mask(SEED, 18, 0.0, 0, 8, 768)
    assert bool((ones == 1).all())


def test_encoder_forward_backward_with_dropout_matches_oracle_under_the_same_masks(fixed_seed):
    from xlxmert_b200.encoder import B200LxmertEncoder
    d = DD
    B, L, V = 3, 20, 64
    sd = P.init_state_dict(P.model_param_specs(d), seed=4, randomize_ln_bias=True)
    batch = synth.make_batch(d, B, L, V, seed=6)
    feats = synth.visual_feats_from(synth.centroid_table(d), batch["cluster_ids"])
    with O.dropout_masks(None):
        emb = O.embeddings(O.sub(sd, "embeddings"), batch["input_ids"])
    mask = O.extended_mask(batch["attention_mask"], torch.float32)
    sdo = {k: v.clone().requires_grad_(True) for k, v in O.sub(sd, "encoder").items()}
    emb_o, feats_o = emb.clone().requires_grad_(True), feats.clone().requires_grad_(True)
    with O.dropout_masks(provider):
        ls, vs = O.encoder(sdo, emb_o, mask, feats_o, batch["visual_pos"], None, heads=d.heads, n_l=d.l_layers,
                           n_r=d.r_layers, n_x=d.x_layers)
    pl, pv = probes([ls[-1].shape, vs[-1].shape], seed=83)
    ((ls[-1] * pl).sum() + (vs[-1] * pv).sum()).backward()

    enc = B200LxmertEncoder(dims=d)
    enc.load_state_dict(O.sub(sd, "encoder"), strict=True)
    enc = enc.cuda().train()
    emb_g, feats_g = emb.cuda().requires_grad_(True), feats.cuda().requires_grad_(True)
    (v, _), (l, _), _ = enc(emb_g, mask.cuda(), feats_g, batch["visual_pos"].cuda())
    assert rel_err(l[-1].detach().cpu(), ls[-1].detach()) < 1e-4
    assert rel_err(v[-1].detach().cpu(), vs[-1].detach()) < 1e-4
    ((l[-1] * pl.cuda()).sum() + (v[-1] * pv.cuda()).sum()).backward()
    assert rel_err(emb_g.grad.cpu(), emb_o.grad) < 1e-3
    assert rel_err(feats_g.grad.cpu(), feats_o.grad) < 1e-3
    for name, p in enc.named_parameters():
        ref = sdo[name].grad
        if name.endswith("key.bias"):
            continue                      # mathematically zero with or without dropout on the probabilities
        assert rel_err(p.grad.cpu(), ref) < 1e-3, (name, rel_err(p.grad.cpu(), ref))
    # the masks really took part: the eval-mode forward differs, and equals the p = 0 module bit for bit
    enc.eval()
    with torch.no_grad():
        (v0, _), (l0, _), _ = enc(emb.cuda(), mask.cuda(), feats.cuda(), batch["visual_pos"].cuda())
    assert rel_err(l0[-1].cpu(), ls[-1].detach()) > 1e-2
    plain = B200LxmertEncoder(dims=replace(d, hidden_dropout=0.0, attention_dropout=0.0))
    plain.load_state_dict(O.sub(sd, "encoder"), strict=True)
    plain = plain.cuda().train()
    with torch.no_grad():
        (v1, _), (l1, _), _ = plain(emb.cuda(), mask.cuda(), feats.cuda(), batch["visual_pos"].cuda())
    assert torch.equal(l0[-1], l1[-1]) and torch.equal(v0[-1], v1[-1])


def test_model_with_embedding_dropout_matches_oracle_under_the_same_masks(fixed_seed):
    """LxmertModel level: adds the embeddings site (HF:212) and the pooler on top of the dropped language output."""
    from xlxmert_b200.lxmert import B200LxmertModel
    d = DD
    B, L, V = 2, 13, 36
    sd = P.init_state_dict(P.model_param_specs(d), seed=8, randomize_ln_bias=True)
    batch = synth.make_batch(d, B, L, V, seed=2)
    feats = synth.visual_feats_from(synth.centroid_table(d), batch["cluster_ids"])
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with O.dropout_masks(provider):
        lang, vis, pooled, _, _ = O.lxmert_model(sdo, batch["input_ids"], feats, batch["visual_pos"],
                                                 batch["attention_mask"], heads=d.heads, n_l=d.l_layers, n_r=d.r_layers,
                                                 n_x=d.x_layers)
    pl, pv, pp = probes([lang.shape, vis.shape, pooled.shape], seed=3)
    ((lang * pl).sum() + (vis * pv).sum() + (pooled * pp).sum()).backward()
    model = B200LxmertModel(d)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model(input_ids=batch["input_ids"].cuda(), visual_feats=feats.cuda(), visual_pos=batch["visual_pos"].cuda(),
                attention_mask=batch["attention_mask"].cuda())
    assert rel_err(out[0].detach().cpu(), lang.detach()) < 1e-4
    assert rel_err(out[1].detach().cpu(), vis.detach()) < 1e-4
    assert rel_err(out[2].detach().cpu(), pooled.detach()) < 1e-4
    ((out[0] * pl.cuda()).sum() + (out[1] * pv.cuda()).sum() + (out[2] * pp.cuda()).sum()).backward()
    for name in ("embeddings.word_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.position_embeddings.weight",
                 "pooler.dense.weight", "encoder.layer.0.attention.self.query.weight",
                 "encoder.x_layers.1.visn_output.dense.weight", "encoder.visn_fc.visn_fc.weight"):
        got, ref = dict(model.named_parameters())[name].grad.cpu(), sdo[name].grad
        assert rel_err(got, ref) < 1e-3, (name, rel_err(got, ref))


def test_fresh_seed_every_step_and_hf_config_probabilities():
    """Two training forwards draw different masks (seed from torch's CPU generator, reproducible under manual_seed);
    dims taken from an HF config carry its 0.1 / 0.1."""
    from transformers import LxmertConfig
    from xlxmert_b200.encoder import B200LxmertEncoder, dims_from_hf_config
    dims = dims_from_hf_config(LxmertConfig())
    assert dims.hidden_dropout == pytest.approx(0.1) and dims.attention_dropout == pytest.approx(0.1)
    d = DD
    sd = P.init_state_dict(P.model_param_specs(d), seed=4)
    enc = B200LxmertEncoder(dims=d)
    enc.load_state_dict(O.sub(sd, "encoder"), strict=True)
    enc = enc.cuda().train()
    batch = synth.make_batch(d, 2, 20, 64, seed=1)
    args = (torch.randn(2, 20, d.hidden).cuda(), None, torch.rand(2, 64, d.feat_dim).cuda(), batch["visual_pos"].cuda())

    def run():
        with torch.no_grad():
            (v, _), (l, _), _ = enc(*args)
        return l[-1]
    torch.manual_seed(7)
    a, b = run(), run()
    torch.manual_seed(7)
    c = run()
    assert not torch.equal(a, b) and torch.equal(a, c)
