"""GPU parity of the two remaining pre-training losses and of the heads' differentiable ``forward``:

* the feature-regression loss (``lxrt/modeling.py:270-284``, the published default ``--visualLosses obj,feat``) and the
  answer head / QA loss (``--taskQA``, ``modeling.py:89-90,286-299``) against the golden vector the reference's own
  ``XLxmertForPretraining`` produced (``oracle/make_golden.py::golden_pretrain_feat_qa``);
* the same losses against the oracle on tiny dimensions for the cases the golden does not hold (features only, masks that
  do not coincide with the labels, a sample without a masked cell, every QA label ignored);
* ``head(hidden)`` with a caller-side loss (the fine-tune pattern, ``tasks/vqa_model.py`` + BCE) against autograd
  through the oracle.

Losses 1e-4 relative, gradients 1e-3 relative (BASELINE.json north_star), arg-max margin-stratified."""
import pytest
import torch

from oracle import lxrt_oracle as O
from xlxmert_b200 import params as P
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D, TINY_DIMS

from util import load_golden, rel_err

pytestmark = pytest.mark.gpu

WATCH = {"ans0w": "answer_head.logit_fc.0.weight", "ans0b": "answer_head.logit_fc.0.bias",
         "ans2w": "answer_head.logit_fc.2.weight", "ans2b": "answer_head.logit_fc.2.bias",
         "ans3w": "answer_head.logit_fc.3.weight", "ans3b": "answer_head.logit_fc.3.bias",
         "poolw": "bert.pooler.dense.weight", "featw": "obj_predict_head.linear_feat.weight",
         "featb": "obj_predict_head.linear_feat.bias", "objtw": "obj_predict_head.transform.dense.weight",
         "clsb": "obj_predict_head.out_cluster.bias", "visnw": "bert.encoder.visn_fc.visn_fc.weight",
         "maskf": "mask_feat", "l0q": "bert.encoder.layer.0.attention.self.query.weight"}


@pytest.fixture(scope="module")
def setup():
    from test_pretrain_parity import _cls_specs
    from xlxmert_b200.pretraining import B200XLxmertForPretraining
    g = load_golden("pretrain_feat_qa_b3")
    B, L, V, wseed, bseed, n_answers = (int(x) for x in g["meta"])
    sd_bert = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(D), seed=wseed + 1, randomize_ln_bias=True)
    sd_cls = P.init_state_dict(_cls_specs(D), seed=wseed + 2, randomize_ln_bias=True)
    sd_ans = P.init_state_dict(P.answerhead_param_specs(D, n_answers), seed=wseed + 4, randomize_ln_bias=True)
    table = synth.centroid_table(D)
    mask_feat = 0.05 * torch.randn(D.feat_dim, generator=torch.Generator().manual_seed(wseed + 3))
    model = B200XLxmertForPretraining(D, num_clusters=D.num_clusters, visual_losses=("obj", "feat"), task_qa=True,
                                      num_qa_labels=n_answers)
    model.set_visual_embedding(table.clone())
    full = {"bert." + k: v for k, v in sd_bert.items()}
    full.update({"obj_predict_head." + k: v for k, v in sd_head.items() if k != "out_cluster.weight"})
    full.update({"cls." + k: v for k, v in sd_cls.items()})
    full.update({"answer_head." + k: v for k, v in sd_ans.items()})
    full["mask_feat"] = mask_feat
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    assert all(any(s in k for s in ("vis_emb", "out_cluster.weight", "decoder.weight")) for k in missing), missing
    # state-dict contract of the answer head: HF's names
    assert {k for k in model.state_dict() if k.startswith("answer_head.")} == {"answer_head." + k for k in sd_ans}
    batch = {k: v.cuda() for k, v in synth.make_batch(D, B, L, V, seed=bseed).items()}
    feat_labels, _ = synth.feat_qa_targets(D, B, bseed, n_answers)
    return g, model.cuda().train(), batch, feat_labels.cuda()


@pytest.mark.parametrize("task", ["vis_mask", "word_mask", "matched"])
def test_feat_and_qa_losses_match_reference_golden(setup, task):
    g, model, batch, feat_labels = setup
    ids = batch["masked_input_ids"] if task == "word_mask" else batch["input_ids"]
    labels = dict(word_labels=batch["word_labels"], obj_labels=batch["obj_labels"],
                  matched_labels=batch["matched_labels"], feat_labels=feat_labels,
                  qa_labels=torch.as_tensor(g[f"qa_labels_{task}"]).cuda())
    model.zero_grad(set_to_none=True)
    out = model(input_ids=ids, visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
                cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"], token_type_ids=batch["token_type_ids"],
                label_dict=labels, task=task)
    out["total_loss"].backward()
    keys = {"vis_mask": ["obj_loss", "feat_loss", "vis_loss"], "word_mask": ["lm_loss"], "matched": ["matched_loss"]}[task]
    for key in keys + ["qa_loss", "total_loss"]:
        ref = float(g[f"{key}_{task}"])
        assert abs(float(out[key].detach()) - ref) < 1e-4 * abs(ref), (key, float(out[key].detach()), ref)
        assert key == "total_loss" or not out[key].requires_grad
    assert set(out) == set(keys) | {"qa_loss", "qa_pred", "total_loss"}
    # answer arg-max: exact where the reference's own top-2 margin is resolvable
    score = torch.as_tensor(g["qa_score"])
    if task != "word_mask":       # qa_score was recorded on the unmasked caption
        top2 = score.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 2e-4
        assert torch.equal(out["qa_pred"].cpu()[safe], torch.as_tensor(g[f"qa_pred_{task}"])[safe])
    assert out["qa_pred"].dtype == torch.int64 and not out["qa_pred"].requires_grad
    params = dict(model.named_parameters())
    for short, name in WATCH.items():
        ref_n = float(g[f"gnorm_{short}_{task}"])
        gr = params[name].grad
        if ref_n != ref_n:                      # NaN marks "no gradient" in the golden
            assert gr is None, (task, name)
            continue
        assert gr is not None, (task, name)
        assert abs(float(gr.norm()) - ref_n) <= 2e-3 * ref_n + 1e-12, (task, name, float(gr.norm()), ref_n)
        if ref_n > 0:
            assert rel_err(gr.flatten()[:16].cpu(), g[f"ghead_{short}_{task}"]) < 2e-3, (task, name)


def _tiny_obj_head(seed=4):
    from xlxmert_b200.heads import B200LxmertVisualObjHead
    d = TINY_DIMS
    sdh = P.init_state_dict(P.objhead_param_specs(d), seed=seed, randomize_ln_bias=True)
    head = B200LxmertVisualObjHead(d, d.num_clusters)
    head.load_state_dict(sdh, strict=True)
    return d, sdh, head.cuda()


def _oracle_visual_losses(sdh, h, obj_labels, feat_labels, vis_mask):
    sdo = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    ho = h.clone().requires_grad_(True)
    feat, logits = O.obj_head(sdo, ho)
    res = {}
    if obj_labels is not None:
        res["obj"] = O.cross_entropy_mean(logits.reshape(-1, logits.shape[-1]), obj_labels.reshape(-1))
    if feat_labels is not None:
        res["feat"] = O.feat_loss(feat, feat_labels, vis_mask)
    return sdo, ho, res


@pytest.mark.parametrize("compact", [True, False])
@pytest.mark.parametrize("case", ["obj+feat", "feat_only", "disjoint", "empty_sample"])
def test_visual_losses_match_oracle(case, compact):
    d, sdh, head = _tiny_obj_head()
    head.compact_rows = compact
    B, V = 3, 36
    g = torch.Generator().manual_seed(11)
    h = torch.randn(B, V, d.hidden, generator=g)
    vis_mask = torch.rand(B, V, generator=g) < 0.4
    feat_labels = torch.randn(B, V, d.feat_dim, generator=g)
    obj_labels = torch.randint(0, d.num_clusters, (B, V), generator=g)
    obj_labels[~vis_mask] = -100                                   # the trainer's rule (lxmert_pretrain.py:163-165)
    if case == "feat_only":
        obj_labels = None
    elif case == "disjoint":                                       # labels and mask need not coincide at this boundary
        obj_labels = torch.randint(0, d.num_clusters, (B, V), generator=g)
        obj_labels[torch.rand(B, V, generator=g) < 0.5] = -100
    elif case == "empty_sample":
        vis_mask[1] = False                                        # n_mask.clamp(min=1) (modeling.py:281)
        obj_labels[1] = -100
    sdo, ho, want = _oracle_visual_losses(sdh, h, obj_labels, feat_labels, vis_mask)
    sum(want.values()).backward()
    hg = h.cuda().requires_grad_(True)
    got = head.losses(hg, obj_labels=None if obj_labels is None else obj_labels.cuda(),
                      feat_labels=feat_labels.cuda(), vis_mask=vis_mask.cuda())
    assert set(got) == set(want)
    head.zero_grad(set_to_none=True)
    sum(got.values()).backward()
    for k in want:
        assert abs(float(got[k].detach()) - float(want[k].detach())) < 1e-4 * abs(float(want[k].detach())), k
    assert rel_err(hg.grad.cpu(), ho.grad) < 1e-3
    for name, p in head.named_parameters():
        ref = sdo[name].grad
        if name == "out_cluster.weight":
            assert p.grad is None                                  # the frozen centroid table (modeling.py:146-151)
        elif ref is None:
            assert p.grad is None, name                            # features only: the classifier is outside the graph
        else:
            assert rel_err(p.grad.cpu(), ref) < 1e-3, name


def test_cluster_head_forward_is_differentiable():
    """``head(hidden, out_keys)`` (modeling.py:38-53) with a caller-side loss on both outputs."""
    d, sdh, head = _tiny_obj_head(seed=6)
    g = torch.Generator().manual_seed(2)
    h = torch.randn(2, 19, d.hidden, generator=g)
    wl = torch.randn(2, 19, d.num_clusters, generator=g)
    wf = torch.randn(2, 19, d.feat_dim, generator=g)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    ho = h.clone().requires_grad_(True)
    feat_o, logits_o = O.obj_head(sdo, ho)
    ((logits_o * wl).sum() + (feat_o * wf).sum()).backward()
    hg = h.cuda().requires_grad_(True)
    out = head(hg, out_keys=["feat", "obj"])
    assert out["obj"].requires_grad and out["feat"].requires_grad
    assert rel_err(out["obj"].detach().cpu(), logits_o.detach()) < 1e-4
    ((out["obj"] * wl.cuda()).sum() + (out["feat"] * wf.cuda()).sum()).backward()
    assert rel_err(hg.grad.cpu(), ho.grad) < 1e-3
    for name, p in head.named_parameters():
        if name != "out_cluster.weight":
            assert rel_err(p.grad.cpu(), sdo[name].grad) < 1e-3, name
    # one output only: the other branch contributes nothing
    head.zero_grad(set_to_none=True)
    hg2 = h.cuda().requires_grad_(True)
    (head(hg2, out_keys=["feat"])["feat"] * wf.cuda()).sum().backward()
    ho2 = h.clone().requires_grad_(True)
    (O.obj_head(sdh, ho2)[0] * wf).sum().backward()
    assert rel_err(hg2.grad.cpu(), ho2.grad) < 1e-3
    assert head.out_cluster.bias.grad is None
    # under no_grad the sampler's fast path is taken and nothing is recorded
    with torch.no_grad():
        assert not head(h.cuda(), out_keys=["obj"])["obj"].requires_grad


def test_answer_head_forward_backward_and_loss():
    """HF ``LxmertVisualAnswerHead`` with a fine-tune style loss (tasks/vqa.py: BCE-with-logits on soft targets) and the
    fused cross-entropy; answer count not a multiple of 8."""
    from xlxmert_b200.heads import B200LxmertVisualAnswerHead
    d, n_answers, B = TINY_DIMS, 91, 37
    sda = P.init_state_dict(P.answerhead_param_specs(d, n_answers), seed=8, randomize_ln_bias=True)
    head = B200LxmertVisualAnswerHead(d, n_answers)
    head.load_state_dict(sda, strict=True)
    head = head.cuda()
    g = torch.Generator().manual_seed(5)
    pooled = torch.tanh(torch.randn(B, d.hidden, generator=g))
    target = torch.rand(B, n_answers, generator=g)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sda.items()}
    po = pooled.clone().requires_grad_(True)
    score_o = O.answer_head(sdo, po)
    torch.nn.functional.binary_cross_entropy_with_logits(score_o, target).backward()
    pg = pooled.cuda().requires_grad_(True)
    score = head(pg)
    assert score.shape == (B, n_answers) and rel_err(score.detach().cpu(), score_o.detach()) < 1e-4
    torch.nn.functional.binary_cross_entropy_with_logits(score, target.cuda()).backward()
    assert rel_err(pg.grad.cpu(), po.grad) < 1e-3
    for name, p in head.named_parameters():
        assert rel_err(p.grad.cpu(), sdo[name].grad) < 1e-3, name
    # fused loss + arg-max, one ignored row
    labels = torch.randint(0, n_answers, (B,), generator=g)
    labels[3] = -100
    sdo = {k: v.clone().requires_grad_(True) for k, v in sda.items()}
    po = pooled.clone().requires_grad_(True)
    loss_o, pred_o = O.qa_loss(sdo, po, labels)
    loss_o.backward()
    head.zero_grad(set_to_none=True)
    pg = pooled.cuda().requires_grad_(True)
    loss, pred = head.loss(pg, labels.cuda())
    loss.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    top2 = O.answer_head(sda, pooled).topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2e-4
    assert torch.equal(pred.cpu()[safe], pred_o[safe])
    assert rel_err(pg.grad.cpu(), po.grad) < 1e-3
    for name, p in head.named_parameters():
        assert rel_err(p.grad.cpu(), sdo[name].grad) < 1e-3, name
    # every label ignored (all pairs flipped in a `matched` step, lxmert_pretrain.py:186-188): NaN like the reference
    nan_loss, _ = head.loss(pooled.cuda(), torch.full((B,), -100, dtype=torch.int64).cuda())
    assert torch.isnan(nan_loss)


def test_pretraining_heads_forward_is_differentiable():
    """HF ``LxmertPreTrainingHeads.forward`` (HF:662-665) → (scores, relationship) with caller-side losses."""
    from xlxmert_b200.heads import B200LxmertPreTrainingHeads
    d = TINY_DIMS
    H = d.hidden
    specs = [("predictions.transform.dense.weight", (H, H)), ("predictions.transform.dense.bias", (H,)),
             ("predictions.transform.LayerNorm.weight", (H,)), ("predictions.transform.LayerNorm.bias", (H,)),
             ("predictions.bias", (d.vocab,)), ("seq_relationship.weight", (2, H)), ("seq_relationship.bias", (2,))]
    sd = P.init_state_dict(specs, seed=3, randomize_ln_bias=True)
    g = torch.Generator().manual_seed(9)
    emb0 = 0.02 * torch.randn(d.vocab, H, generator=g)
    heads = B200LxmertPreTrainingHeads(d, torch.nn.Parameter(emb0.clone()))
    heads.load_state_dict(sd, strict=False)
    heads = heads.cuda()
    lang = torch.randn(2, 7, H, generator=g)
    pooled = torch.tanh(torch.randn(2, H, generator=g))
    ws, wr = torch.randn(2, 7, d.vocab, generator=g), torch.randn(2, 2, generator=g)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sdo["predictions.decoder.weight"] = emb0.clone().requires_grad_(True)
    lo, po = lang.clone().requires_grad_(True), pooled.clone().requires_grad_(True)
    scores_o, rel_o = O.lm_head(sdo, lo, po)
    ((scores_o * ws).sum() + (rel_o * wr).sum()).backward()
    lg, pg = lang.cuda().requires_grad_(True), pooled.cuda().requires_grad_(True)
    scores, rel = heads(lg, pg)
    assert rel_err(scores.detach().cpu(), scores_o.detach()) < 1e-4 and rel_err(rel.detach().cpu(), rel_o.detach()) < 1e-4
    ((scores * ws.cuda()).sum() + (rel * wr.cuda()).sum()).backward()
    assert rel_err(lg.grad.cpu(), lo.grad) < 1e-3 and rel_err(pg.grad.cpu(), po.grad) < 1e-3
    named = dict(heads.named_parameters())
    for name, ref in sdo.items():
        assert rel_err(named[name].grad.cpu(), ref.grad) < 1e-3, name
