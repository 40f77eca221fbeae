"""``accelerate(model)`` — the documented drop-in (SURVEY.md §8b "Recommended interposition"): swap
the HF ``LxmertEncoder`` inside an existing model for the native one *sharing the same ``nn.Parameter`` objects*.

CPU half: the swap on the reference's own ``XLxmertForPretraining`` (imported through oracle/refshim.py when
/root/reference exists) and on a bare HF ``LxmertModel`` keeps every state-dict key and every Parameter object.
GPU half: an accelerated HF ``LxmertModel`` gives the outputs and gradients of the unswapped HF module on CPU."""
import copy

import pytest
import torch

from util import probes, rel_err


def _hf_model(seed=0, **kw):
    from transformers import LxmertConfig, LxmertModel
    torch.manual_seed(seed)
    cfg = LxmertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)
    return LxmertModel(cfg)


def _check_swap(model, encoder_path):
    from xlxmert_b200.encoder import accelerate
    before = {k: v for k, v in model.named_parameters()}
    keys = list(model.state_dict().keys())
    out = accelerate(model)
    assert out is model
    holder = model
    for a in encoder_path[:-1]:
        holder = getattr(holder, a)
    enc = getattr(holder, encoder_path[-1])
    assert type(enc).__name__ == "B200LxmertEncoder"
    assert list(model.state_dict().keys()) == keys                       # checkpoints load / save unchanged
    after = {k: v for k, v in model.named_parameters()}
    assert after.keys() == before.keys()
    assert all(after[k] is before[k] for k in before)                   # optimiser / DDP keep seeing the same objects
    return enc


def test_swap_keeps_keys_and_parameter_objects_on_hf_model():
    enc = _check_swap(_hf_model(l_layers=2, r_layers=1, x_layers=1), ("encoder",))
    assert (enc.dims.l_layers, enc.dims.r_layers, enc.dims.x_layers) == (2, 1, 1)
    # CPU tensors: the accelerated module refuses loudly instead of falling back
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 4, 768), None, torch.zeros(1, 4, 2048), torch.zeros(1, 4, 4))


def test_swap_inside_the_reference_pretraining_model():
    from oracle import refshim
    if not refshim.available():
        pytest.skip("needs the reference checkout (/root/reference), present in the build container only")
    model = refshim.build_pretraining_model(num_clusters=64, l_layers=1, r_layers=1, x_layers=1,
                                            hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    enc = _check_swap(model, ("bert", "encoder"))
    assert enc.dims.hidden_dropout == pytest.approx(0.1)                 # training-mode dropout follows the HF config
    assert model.bert.encoder is enc and model.obj_predict_head is not None


@pytest.mark.gpu
def test_accelerated_hf_model_matches_unswapped_hf_on_cpu():
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.encoder import accelerate
    ref = _hf_model(seed=3).train()
    fast = accelerate(copy.deepcopy(ref).cuda()).train()
    B, L, V = 3, 20, 64
    batch = synth.make_batch(D, B, L, V, seed=5)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"])
    kw = dict(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
              attention_mask=batch["attention_mask"], return_dict=True)
    o = ref(**kw)
    pl, pv, pp = probes([o.language_output.shape, o.vision_output.shape, o.pooled_output.shape], seed=11)
    ((o.language_output * pl).sum() + (o.vision_output * pv).sum() + (o.pooled_output * pp).sum()).backward()
    g = fast(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()})
    assert rel_err(g.language_output.detach().cpu(), o.language_output.detach()) < 1e-4
    assert rel_err(g.vision_output.detach().cpu(), o.vision_output.detach()) < 1e-4
    assert rel_err(g.pooled_output.detach().cpu(), o.pooled_output.detach()) < 1e-4
    ((g.language_output * pl.cuda()).sum() + (g.vision_output * pv.cuda()).sum() + (g.pooled_output * pp.cuda()).sum()).backward()
    refg = dict(ref.named_parameters())
    checked = 0
    for name, p in fast.named_parameters():
        r = refg[name].grad
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        if name.endswith("key.bias"):
            continue
        assert p.grad is not None, name
        assert rel_err(p.grad.cpu(), r) < 1e-3, (name, rel_err(p.grad.cpu(), r))
        checked += 1
    assert checked > 300
    # every layer's hidden state, like HF's encoder, when the swap is asked to keep them (HF's LxmertModel does not tell
    # its encoder whether the caller wants them: accelerate(..., output_hidden_states=True) does)
    full = accelerate(copy.deepcopy(ref).cuda(), output_hidden_states=True).eval()
    with torch.no_grad():
        h = full(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}, output_hidden_states=True)
        hr = ref.eval()(**kw, output_hidden_states=True)
    assert len(h.language_hidden_states) == 14 and len(h.vision_hidden_states) == 10
    for a, b in zip(h.language_hidden_states + h.vision_hidden_states, hr.language_hidden_states + hr.vision_hidden_states):
        assert rel_err(a.cpu(), b) < 1e-4


@pytest.mark.gpu
def test_output_attentions_match_hf():
    """``output_attentions=True`` (a1 / a13 boundary): language_attentions (9), vision_attentions (5) and
    cross_encoder_attentions (5, language queries over vision keys) equal HF's in eval mode — through an accelerated HF
    ``LxmertModel`` and through ``B200LxmertModel``."""
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.encoder import accelerate, dims_from_hf_config
    from xlxmert_b200.lxmert import B200LxmertModel
    ref = _hf_model(seed=4).eval()
    fast = accelerate(copy.deepcopy(ref).cuda()).eval()
    B, L, V = 2, 13, 36
    batch = synth.make_batch(D, B, L, V, seed=8)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"])
    kw = dict(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
              attention_mask=batch["attention_mask"], return_dict=True, output_attentions=True)
    with torch.no_grad():
        o = ref(**kw)
        g = fast(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()})
        native = B200LxmertModel(dims_from_hf_config(ref.config), source=copy.deepcopy(ref).cuda()).eval()
        n = native(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()})
    for got in (g, n):
        assert len(got.language_attentions) == 9 and len(got.vision_attentions) == 5 and len(got.cross_encoder_attentions) == 5
        for a, b in zip(got.language_attentions + got.vision_attentions + got.cross_encoder_attentions,
                        o.language_attentions + o.vision_attentions + o.cross_encoder_attentions):
            assert a.shape == b.shape
            assert float((a.cpu() - b).abs().max()) < 1e-5          # probabilities in [0, 1]: absolute error
        assert rel_err(got.language_output.cpu() if hasattr(got, "language_output") else got[0].cpu(), o.language_output) < 1e-4
    # without the flag nothing is exported (and the small inference plan is used)
    with torch.no_grad():
        plain = native(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items() if k != "output_attentions"})
    assert plain.language_attentions is None and torch.equal(plain[0], n[0])


@pytest.mark.gpu
def test_inputs_embeds_path_matches_hf():
    """``inputs_embeds`` instead of ``input_ids`` (HF:199-203,737-742): same outputs as the id path bit for bit when the
    embeddings are the table rows, and the gradient with respect to ``inputs_embeds`` equals HF's."""
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.encoder import dims_from_hf_config
    from xlxmert_b200.lxmert import B200LxmertModel
    ref = _hf_model(seed=6, l_layers=2, r_layers=1, x_layers=1).train()
    native = B200LxmertModel(dims_from_hf_config(ref.config), source=copy.deepcopy(ref).cuda()).train()
    B, L, V = 2, 20, 64
    batch = synth.make_batch(D, B, L, V, seed=4)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"])
    ids = batch["input_ids"]
    e_ref = ref.embeddings.word_embeddings.weight[ids].detach().clone().requires_grad_(True)
    o = ref(inputs_embeds=e_ref, visual_feats=feats, visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
            return_dict=True)
    pl, pv = probes([o.language_output.shape, o.vision_output.shape], seed=2)
    ((o.language_output * pl).sum() + (o.vision_output * pv).sum()).backward()
    e_gpu = e_ref.detach().clone().cuda().requires_grad_(True)
    kw = dict(visual_feats=feats.cuda(), visual_pos=batch["visual_pos"].cuda(), attention_mask=batch["attention_mask"].cuda())
    g = native(inputs_embeds=e_gpu, **kw)
    assert rel_err(g[0].detach().cpu(), o.language_output.detach()) < 1e-4
    ((g[0] * pl.cuda()).sum() + (g[1] * pv.cuda()).sum()).backward()
    assert rel_err(e_gpu.grad.cpu(), e_ref.grad) < 1e-3
    assert native.embeddings.word_embeddings.weight.grad is None           # the table took no part
    assert rel_err(native.embeddings.position_embeddings.weight.grad.cpu(), ref.embeddings.position_embeddings.weight.grad) < 1e-3
    with torch.no_grad():
        by_id = native(input_ids=ids.cuda(), **kw)
        by_emb = native(inputs_embeds=e_gpu.detach(), **kw)
    assert torch.equal(by_id[0], by_emb[0]) and torch.equal(by_id[1], by_emb[1])
    with pytest.raises(ValueError):
        native(input_ids=ids.cuda(), inputs_embeds=e_gpu.detach(), **kw)
