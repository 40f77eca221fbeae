"""The oracle (oracle/*.py) against golden vectors produced by the reference's own classes
(oracle/make_golden.py: HF LxmertModel, lxrt.modeling.XLxmertForPretraining, layers.Generator)."""
import numpy as np
import pytest
import torch

from oracle import generator_oracle as GO
from oracle import lxrt_oracle as O
from xlxmert_b200 import params as P
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D

from util import checksum, load_golden, probes, rel_err

TOL = 2e-5   # fp32 CPU summation-order noise only: both sides are fp32 on the same inputs


def _model_case(name):
    g = load_golden(name)
    B, L, V, wseed, bseed = (int(x) for x in g["meta"])
    sd = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    assert abs(checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"])), \
        "seeded weights differ from the ones the golden was made with (torch RNG drift?)"
    batch = synth.make_batch(D, B, L, V, seed=bseed)
    feats = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"])
    return g, sd, batch, feats


@pytest.mark.parametrize("name", ["model_b2_l20_v64", "model_b3_l13_v36"])
def test_model_forward_matches_hf_golden(name):
    g, sd, batch, feats = _model_case(name)
    with torch.no_grad():
        lang, vis, pooled, ls, vs = O.lxmert_model(sd, batch["input_ids"], feats, batch["visual_pos"],
                                                    batch["attention_mask"])
    assert rel_err(lang, g["lang"]) < TOL
    assert rel_err(vis, g["vis"]) < TOL
    assert rel_err(pooled, g["pooled"]) < TOL
    assert len(ls) == 14 and len(vs) == 10          # 9 L + 5 X, 5 R + 5 X (HF:524-550)
    for i, h in enumerate(ls):
        assert rel_err(h[:, ::4, ::8], g[f"lang_h{i}"]) < TOL, i
    for i, h in enumerate(vs):
        assert rel_err(h[:, ::8, ::8], g[f"vis_h{i}"]) < TOL, i


def test_model_backward_matches_hf_golden():
    g, sd, batch, feats = _model_case("model_b2_l20_v64")
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feats = feats.clone().requires_grad_(True)
    lang, vis, pooled, _, _ = O.lxmert_model(sd, batch["input_ids"], feats, batch["visual_pos"],
                                             batch["attention_mask"])
    pl, pv, pp = probes([lang.shape, vis.shape, pooled.shape], seed=int(g["meta"][4]) + 77)
    loss = (lang * pl).sum() + (vis * pv).sum() + (pooled * pp).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"])) + 1e-3
    assert rel_err(feats.grad[:, ::8, ::64], g["dfeats_sub"]) < 1e-4
    for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        gr = sd[str(name)].grad
        assert gr is not None, name
        if norm < 1e-5:      # mathematically-zero grads (key bias: softmax is shift-invariant) are fp32 noise
            assert float(gr.double().norm()) < 1e-5, name
            continue
        assert abs(float(gr.double().norm()) - norm) <= 1e-4 * norm, name
        n = min(8, gr.numel())
        assert np.allclose(gr.flatten()[:n].double().numpy(), head[:n], rtol=2e-3, atol=1e-5 * max(norm, 1e-3)), name


def _pretrain_case():
    g = load_golden("pretrain_b2")
    B, L, V, wseed, bseed = (int(x) for x in g["meta"])
    sd_bert = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(D), seed=wseed + 1, randomize_ln_bias=True)
    H = D.hidden
    cls_specs = [("predictions.transform.dense.weight", (H, H)), ("predictions.transform.dense.bias", (H,)),
                 ("predictions.transform.LayerNorm.weight", (H,)), ("predictions.transform.LayerNorm.bias", (H,)),
                 ("predictions.bias", (D.vocab,)),
                 ("seq_relationship.weight", (2, H)), ("seq_relationship.bias", (2,))]
    sd_cls = P.init_state_dict(cls_specs, seed=wseed + 2, randomize_ln_bias=True)
    sd_cls["predictions.decoder.weight"] = sd_bert["embeddings.word_embeddings.weight"]
    table = synth.centroid_table(D)
    sd_head["out_cluster.weight"] = table      # tied to the centroid table (modeling.py:146-151)
    mask_feat = 0.05 * torch.randn(D.feat_dim, generator=torch.Generator().manual_seed(wseed + 3))
    batch = synth.make_batch(D, B, L, V, seed=bseed)
    return g, sd_bert, sd_head, sd_cls, table, mask_feat, batch


def test_pretraining_losses_and_cluster_head_match_reference_golden():
    g, sd_bert, sd_head, sd_cls, table, mask_feat, batch = _pretrain_case()
    mask_feat = mask_feat.clone().requires_grad_(True)
    feats = O.mask_visual_feats(table[batch["cluster_ids"]], batch["vis_mask"], mask_feat)
    lang, vis, pooled, _, _ = O.lxmert_model(sd_bert, batch["input_ids"], feats, batch["visual_pos"],
                                             batch["attention_mask"])
    loss = O.obj_loss(sd_head, vis, batch["obj_labels"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss_vis_mask"])) < 1e-5 * float(g["loss_vis_mask"])
    assert rel_err(mask_feat.grad[:16], g["grad_mask_feat_head"]) < 1e-3
    with torch.no_grad():
        feat, logits = O.obj_head(sd_head, vis)
        prob, idx = O.sampler_predict(logits)
        assert rel_err(feat[:, ::8, ::32], g["head_feat_sub"]) < TOL
        assert rel_err(logits[:, ::8, ::100], g["head_logits_sub"]) < TOL
        # cluster argmax: exact wherever the reference's own top-2 margin exceeds fp32 noise
        safe = torch.as_tensor(g["head_margin"]) > 1e-4
        assert torch.equal(idx[safe], torch.as_tensor(g["head_argmax"])[safe])
        assert rel_err(prob, g["head_maxprob"]) < 1e-4

        # word_mask / matched (modeling.py:216-235)
        feats0 = table[batch["cluster_ids"]]
        lang, vis, pooled, _, _ = O.lxmert_model(sd_bert, batch["masked_input_ids"], feats0,
                                                 batch["visual_pos"], batch["attention_mask"])
        scores, _ = O.lm_head(sd_cls, lang, pooled)
        lm = O.cross_entropy_mean(scores.view(-1, D.vocab), batch["word_labels"].view(-1))
        assert abs(float(lm) - float(g["loss_word_mask"])) < 1e-5 * float(g["loss_word_mask"])
        lang, vis, pooled, _, _ = O.lxmert_model(sd_bert, batch["input_ids"], feats0,
                                                 batch["visual_pos"], batch["attention_mask"])
        _, rel = O.lm_head(sd_cls, lang, pooled)
        ml = O.cross_entropy_mean(rel, batch["matched_labels"])
        assert abs(float(ml) - float(g["loss_matched"])) < 1e-5 * float(g["loss_matched"])


def test_teacher_forced_nar_sampler_matches_reference_golden():
    g, sd_bert, sd_head, _, table, mask_feat, batch = _pretrain_case()
    B = batch["input_ids"].shape[0]
    code = torch.zeros(B, 64, D.feat_dim)
    with torch.no_grad():
        for i in range(4):
            vis_mask = torch.as_tensor(g[f"nar_mask{i}"])
            assert int(vis_mask[0].sum()) == O.nar_n_mask(i, 4)
            code, prob, idx = O.sampler_step(sd_bert, sd_head, table, mask_feat, batch["input_ids"],
                                             batch["visual_pos"], code, vis_mask)
            assert rel_err(prob, g[f"nar_prob{i}"]) < 1e-4, i
            same = (idx == torch.as_tensor(g[f"nar_id{i}"])).float().mean()
            assert same == 1.0, (i, float(same))
        assert rel_err(code[:, :, ::64], g["nar_code_sub"]) < TOL


def test_bilinear_restated_equals_aten():
    x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(3))
    for size in (8, 16, 64, 256):
        ref = torch.nn.functional.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
        assert torch.allclose(GO.bilinear_resize(x, size), ref, atol=1e-6)


def test_generator_matches_reference_golden():
    g = load_golden("generator_b2")
    B, wseed, bseed = (int(x) for x in g["meta"])
    sd = P.init_generator_state_dict(seed=wseed)
    assert abs(checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    batch = synth.make_batch(D, B, 20, 64, seed=bseed)
    code = synth.centroid_table(D)[batch["cluster_ids"]]
    emb = code.permute(0, 2, 1).reshape(B, 2048, 8, 8)
    with torch.no_grad():
        img, aux = GO.generator(sd, emb, return_intermediates=True)
    assert float(g["saturated_frac"]) < 0.5, "vacuous parity target: most pixels saturated"
    for i, h in enumerate(aux["h"]):
        s = max(1, h.shape[-1] // 16)
        assert rel_err(h[:, :, ::s, ::s], g[f"h{i}_sub"]) < 1e-4, i
    assert rel_err(aux["pre_tanh"][:, :, ::4, ::4], g["pre_tanh_sub"]) < 1e-4
    assert rel_err(img[:, :, ::4, ::4], g["img_sub"]) < 1e-4


def _feat_qa_case():
    g = load_golden("pretrain_feat_qa_b3")
    B, L, V, wseed, bseed, n_answers = (int(x) for x in g["meta"])
    sd_bert = P.init_state_dict(P.model_param_specs(D), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(D), seed=wseed + 1, randomize_ln_bias=True)
    sd_ans = P.init_state_dict(P.answerhead_param_specs(D, n_answers), seed=wseed + 4, randomize_ln_bias=True)
    table = synth.centroid_table(D)
    sd_head["out_cluster.weight"] = table
    mask_feat = 0.05 * torch.randn(D.feat_dim, generator=torch.Generator().manual_seed(wseed + 3))
    batch = synth.make_batch(D, B, L, V, seed=bseed)
    feat_labels, qa_labels = synth.feat_qa_targets(D, B, bseed, n_answers)
    chk = feat_labels.double().abs().sum().item() + qa_labels.sum().item()
    assert abs(chk - float(g["inputs_checksum"])) < 1e-6 * chk
    return g, sd_bert, sd_head, sd_ans, table, mask_feat, batch, feat_labels


def test_feat_regression_and_qa_losses_match_reference_golden():
    """modeling.py:270-299 as published (obj + feat visual losses, --taskQA), vis_mask task."""
    g, sd_bert, sd_head, sd_ans, table, mask_feat, batch, feat_labels = _feat_qa_case()
    for t in sd_ans.values():
        t.requires_grad_(True)
    sd_head["linear_feat.weight"].requires_grad_(True)
    mask_feat = mask_feat.clone().requires_grad_(True)
    feats = O.mask_visual_feats(table[batch["cluster_ids"]], batch["vis_mask"], mask_feat)
    lang, vis, pooled, _, _ = O.lxmert_model(sd_bert, batch["input_ids"], feats, batch["visual_pos"],
                                             batch["attention_mask"])
    feat, logits = O.obj_head(sd_head, vis)
    obj = O.cross_entropy_mean(logits.reshape(-1, logits.shape[-1]), batch["obj_labels"].reshape(-1))
    fl = O.feat_loss(feat, feat_labels, batch["vis_mask"])
    qa, pred = O.qa_loss(sd_ans, pooled, torch.as_tensor(g["qa_labels_vis_mask"]))
    total = obj + fl + qa
    total.backward()
    for mine, key in ((obj, "obj_loss"), (fl, "feat_loss"), (qa, "qa_loss"), (obj + fl, "vis_loss"), (total, "total_loss")):
        ref = float(g[f"{key}_vis_mask"])
        assert abs(float(mine.detach()) - ref) < 1e-5 * abs(ref), key
    assert torch.equal(pred, torch.as_tensor(g["qa_pred_vis_mask"]))
    for short, t in (("ans0w", sd_ans["logit_fc.0.weight"]), ("ans2b", sd_ans["logit_fc.2.bias"]),
                     ("ans3w", sd_ans["logit_fc.3.weight"]), ("ans3b", sd_ans["logit_fc.3.bias"]),
                     ("featw", sd_head["linear_feat.weight"]), ("maskf", mask_feat)):
        ref_n = float(g[f"gnorm_{short}_vis_mask"])
        assert abs(float(t.grad.norm()) - ref_n) < 1e-3 * ref_n, short
        assert rel_err(t.grad.flatten()[:16], g[f"ghead_{short}_vis_mask"]) < 1e-3, short
    with torch.no_grad():
        feats0 = table[batch["cluster_ids"]]
        _, _, pooled0, _, _ = O.lxmert_model(sd_bert, batch["input_ids"], feats0, batch["visual_pos"],
                                             batch["attention_mask"])
        assert rel_err(pooled0, g["pooled"]) < TOL
        assert rel_err(O.answer_head(sd_ans, pooled0), g["qa_score"]) < TOL
        # matched task: the labels of flipped pairs are ignored (lxmert_pretrain.py:186-188)
        qa_m, _ = O.qa_loss(sd_ans, pooled0, torch.as_tensor(g["qa_labels_matched"]))
        assert abs(float(qa_m) - float(g["qa_loss_matched"])) < 1e-5 * float(g["qa_loss_matched"])
