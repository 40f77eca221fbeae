"""ORACLE tooling — generate ``tests/golden/*.npz`` from the reference's own classes.

Runs only in the build container (needs ``/root/reference`` and the installed HF ``transformers``):

    python -m oracle.make_golden            # writes tests/golden/*.npz

Inputs and weights are *not* stored: they are regenerated at test time from the seeds recorded in
each file by ``xlxmert_b200.synth`` / ``xlxmert_b200.params`` (a checksum of the weights is stored so
that RNG drift is detected rather than misread as a parity failure).  Outputs are stored in full
where small and sub-sampled (fixed strides) where large.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from xlxmert_b200 import params as P  # noqa: E402
from xlxmert_b200 import synth  # noqa: E402
from xlxmert_b200.config import DEFAULT_DIMS  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(sd) -> float:
    """Order-independent fp64 fingerprint of a state dict."""
    return float(sum((v.double().abs().sum() + v.double().sum() * 0.5) for v in sd.values()
                     if v.is_floating_point()))


def probes(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


def hf_model(d, sd):
    from transformers import LxmertConfig, LxmertModel
    cfg = LxmertConfig(hidden_size=d.hidden, num_attention_heads=d.heads, intermediate_size=d.intermediate,
                       visual_feat_dim=d.feat_dim, l_layers=d.l_layers, r_layers=d.r_layers,
                       x_layers=d.x_layers, vocab_size=d.vocab, max_position_embeddings=d.max_pos,
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = LxmertModel(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m.eval()


def golden_model(name, d, B, L, V, wseed, bseed):
    """HF ``LxmertModel`` forward + backward (config 1 of BASELINE.json: the parity anchor)."""
    sd = P.init_state_dict(P.model_param_specs(d), seed=wseed, randomize_ln_bias=True)
    table = synth.centroid_table(d)
    batch = synth.make_batch(d, B, L, V, seed=bseed)
    feats = synth.visual_feats_from(table, batch["cluster_ids"]).clone().requires_grad_(True)
    m = hf_model(d, sd)
    out = m(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
            attention_mask=batch["attention_mask"], output_hidden_states=True, return_dict=True)
    lang, vis, pooled = out.language_output, out.vision_output, out.pooled_output
    pl, pv, pp = probes([lang.shape, vis.shape, pooled.shape], seed=bseed + 77)
    loss = (lang * pl).sum() + (vis * pv).sum() + (pooled * pp).sum()
    loss.backward()
    rec = dict(lang=lang, vis=vis, pooled=pooled, loss=loss.detach(),
               weights_checksum=torch.tensor(checksum(sd), dtype=torch.float64),
               meta=np.array([B, L, V, wseed, bseed]),
               dfeats_sub=feats.grad[:, ::8, ::64])
    for i, h in enumerate(out.language_hidden_states):
        rec[f"lang_h{i}"] = h[:, ::4, ::8]
    for i, h in enumerate(out.vision_hidden_states):
        rec[f"vis_h{i}"] = h[:, ::8, ::8]
    names, norms, heads = [], [], []
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        heads.append(p.grad.flatten()[:8].double().numpy().copy())
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms)
    rec["grad_heads"] = np.stack([np.pad(h, (0, 8 - len(h))) for h in heads])
    save(name, rec)


def golden_pretrain(name, d, B, wseed, bseed):
    """Reference ``XLxmertForPretraining.forward`` per task + cluster head + teacher-forced sampler."""
    model = refshim.build_pretraining_model(
        num_clusters=d.num_clusters, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
        task_qa=False)  # pretrain.bash:24-26 enables MaskLM, ObjPredict, Matched only
    sd_bert = P.init_state_dict(P.model_param_specs(d), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(d), seed=wseed + 1, randomize_ln_bias=True)
    cls_specs = [("predictions.transform.dense.weight", (d.hidden, d.hidden)),
                 ("predictions.transform.dense.bias", (d.hidden,)),
                 ("predictions.transform.LayerNorm.weight", (d.hidden,)),
                 ("predictions.transform.LayerNorm.bias", (d.hidden,)),
                 ("predictions.bias", (d.vocab,)),
                 ("seq_relationship.weight", (2, d.hidden)), ("seq_relationship.bias", (2,))]
    sd_cls = P.init_state_dict(cls_specs, seed=wseed + 2, randomize_ln_bias=True)
    table = synth.centroid_table(d)
    g = torch.Generator().manual_seed(wseed + 3)
    mask_feat = 0.05 * torch.randn(d.feat_dim, generator=g)

    model.set_visual_embedding(table.clone())
    full = {"bert." + k: v for k, v in sd_bert.items()}
    full.update({"obj_predict_head." + k: v for k, v in sd_head.items() if k != "out_cluster.weight"})
    full.update({"cls." + k: v for k, v in sd_cls.items()})
    full["mask_feat"] = mask_feat
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    bad = [k for k in missing if not any(s in k for s in ("position_ids", "vis_emb", "out_cluster.weight",
                                                          "decoder.weight", "decoder.bias"))]
    assert not bad, bad
    model.eval()
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    rec = dict(meta=np.array([B, 20, 64, wseed, bseed]),
               weights_checksum=torch.tensor(checksum(sd_bert) + checksum(sd_head) + checksum(sd_cls),
                                             dtype=torch.float64))
    common = dict(visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
                  cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"],
                  token_type_ids=batch["token_type_ids"], return_dict=True)
    labels = dict(word_labels=batch["word_labels"], obj_labels=batch["obj_labels"],
                  matched_labels=batch["matched_labels"])
    for task, ids in (("vis_mask", batch["input_ids"]), ("word_mask", batch["masked_input_ids"]),
                      ("matched", batch["input_ids"])):
        model.zero_grad()
        out = model(input_ids=ids, label_dict=labels, task=task, **common)
        out["total_loss"].backward()
        rec[f"loss_{task}"] = out["total_loss"].detach()
        rec[f"gradnorm_mask_feat_{task}"] = torch.tensor(
            0.0 if model.mask_feat.grad is None else model.mask_feat.grad.norm().item())
        if task == "vis_mask":
            rec["grad_mask_feat_head"] = model.mask_feat.grad[:16].clone()
            rec["gradnorm_out_cluster_bias"] = model.obj_predict_head.out_cluster.bias.grad.norm()
            rec["gradnorm_linear_feat_w"] = model.obj_predict_head.linear_feat.weight.grad.norm()
            rec["gradnorm_visn_fc_w"] = model.bert.encoder.visn_fc.visn_fc.weight.grad.norm()

    with torch.no_grad():
        feats = synth.visual_feats_from(table, batch["cluster_ids"])
        feats = torch.where(batch["vis_mask"].unsqueeze(-1), mask_feat.view(1, 1, -1), feats)
        o = model.bert(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
                       attention_mask=batch["attention_mask"], return_dict=True)
        head = model.obj_predict_head(o[1], out_keys=["obj", "feat"])
        logits = head["obj"]
        prob, idx = torch.softmax(logits, dim=2).max(dim=2)
        top2v, top2i = logits.topk(2, dim=2)
        rec.update(head_feat_sub=head["feat"][:, ::8, ::32], head_logits_sub=logits[:, ::8, ::100],
                   head_argmax=idx, head_maxprob=prob, head_margin=top2v[..., 0] - top2v[..., 1], head_top2=top2i)

        # teacher-forced NAR sampling, 4 steps (imggen_model.py:199-243): the masks chosen by the
        # reference run are stored and replayed by the tests (topk tie order is implementation-defined).
        n_steps, n_grids = 4, 64
        ids = batch["input_ids"]
        vpos = batch["visual_pos"]
        for i in range(n_steps):
            n_mask = int((n_steps - i) / n_steps * n_grids)
            if i == 0:
                vis_mask = torch.ones(B, n_grids).long()
                code = torch.zeros(B, n_grids, d.feat_dim)
            else:
                _, lowest_arg = pred_prob.topk(n_mask, dim=1, largest=False)
                vis_mask = torch.zeros(B, n_grids).long()
                vis_mask.scatter_(1, lowest_arg, 1)
            code = torch.where(vis_mask.view(B, n_grids, 1).bool(),
                               model.mask_feat.view(1, 1, -1).to(dtype=code.dtype), code)
            lx = model.bert(input_ids=ids, visual_feats=code, visual_pos=vpos, attention_mask=ids > 0,
                            return_dict=True)
            pl = model.obj_predict_head(lx[1], out_keys=["obj"])["obj"]
            pred_prob, pred_id = torch.softmax(pl, dim=2).max(dim=2)
            code = torch.where(vis_mask.view(B, n_grids, 1).bool(), model.vis_emb(pred_id), code)
            t2v, t2i = pl.topk(2, dim=2)
            rec[f"nar_mask{i}"] = vis_mask
            rec[f"nar_prob{i}"] = pred_prob
            rec[f"nar_id{i}"] = pred_id
            rec[f"nar_top2_{i}"] = t2i
            rec[f"nar_margin{i}"] = t2v[..., 0] - t2v[..., 1]
        rec["nar_code_sub"] = code[:, :, ::64]
    save(name, rec)


def golden_pretrain_feat_qa(name, d, B, wseed, bseed, n_answers=157):
    """The reference model AS PUBLISHED: ``visual_losses`` = obj + feat (modeling.py:117-137; the
    default ``--visualLosses obj,feat``, param.py:123) and ``--taskQA`` (modeling.py:89-90,286-299): the QA loss is
    added to every task's loss; for ``matched`` the labels of flipped pairs are ignored (lxmert_pretrain.py:184-189)."""
    model = refshim.build_pretraining_model(
        num_clusters=d.num_clusters, keep_feat_loss=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
        task_qa=True, num_qa_labels=n_answers, visual_attr_loss=False)   # lxmert_pretrain.py:72-74: from --visualLosses
    sd_bert = P.init_state_dict(P.model_param_specs(d), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(d), seed=wseed + 1, randomize_ln_bias=True)
    cls_specs = [("predictions.transform.dense.weight", (d.hidden, d.hidden)),
                 ("predictions.transform.dense.bias", (d.hidden,)),
                 ("predictions.transform.LayerNorm.weight", (d.hidden,)),
                 ("predictions.transform.LayerNorm.bias", (d.hidden,)),
                 ("predictions.bias", (d.vocab,)),
                 ("seq_relationship.weight", (2, d.hidden)), ("seq_relationship.bias", (2,))]
    sd_cls = P.init_state_dict(cls_specs, seed=wseed + 2, randomize_ln_bias=True)
    sd_ans = P.init_state_dict(P.answerhead_param_specs(d, n_answers), seed=wseed + 4, randomize_ln_bias=True)
    table = synth.centroid_table(d)
    g = torch.Generator().manual_seed(wseed + 3)
    mask_feat = 0.05 * torch.randn(d.feat_dim, generator=g)
    model.set_visual_embedding(table.clone())
    full = {"bert." + k: v for k, v in sd_bert.items()}
    full.update({"obj_predict_head." + k: v for k, v in sd_head.items() if k != "out_cluster.weight"})
    full.update({"cls." + k: v for k, v in sd_cls.items()})
    full.update({"answer_head." + k: v for k, v in sd_ans.items()})
    full["mask_feat"] = mask_feat
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    bad = [k for k in missing if not any(s in k for s in ("position_ids", "vis_emb", "out_cluster.weight",
                                                          "decoder.weight", "decoder.bias"))]
    assert not bad, bad
    model.eval()
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    feat_labels, qa_labels = synth.feat_qa_targets(d, B, bseed, n_answers)
    rec = dict(meta=np.array([B, 20, 64, wseed, bseed, n_answers]),
               weights_checksum=torch.tensor(checksum(sd_bert) + checksum(sd_head) + checksum(sd_cls) + checksum(sd_ans),
                                             dtype=torch.float64),
               inputs_checksum=torch.tensor(feat_labels.double().abs().sum().item() + qa_labels.sum().item(),
                                            dtype=torch.float64))
    common = dict(visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
                  cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"],
                  token_type_ids=batch["token_type_ids"], return_dict=True)
    watch = {"ans0w": "answer_head.logit_fc.0.weight", "ans0b": "answer_head.logit_fc.0.bias",
             "ans2w": "answer_head.logit_fc.2.weight", "ans2b": "answer_head.logit_fc.2.bias",
             "ans3w": "answer_head.logit_fc.3.weight", "ans3b": "answer_head.logit_fc.3.bias",
             "poolw": "bert.pooler.dense.weight", "featw": "obj_predict_head.linear_feat.weight",
             "featb": "obj_predict_head.linear_feat.bias", "objtw": "obj_predict_head.transform.dense.weight",
             "clsb": "obj_predict_head.out_cluster.bias", "visnw": "bert.encoder.visn_fc.visn_fc.weight",
             "maskf": "mask_feat", "l0q": "bert.encoder.layer.0.attention.self.query.weight"}
    params = dict(model.named_parameters())
    for task, ids in (("vis_mask", batch["input_ids"]), ("word_mask", batch["masked_input_ids"]),
                      ("matched", batch["input_ids"])):
        qa = qa_labels.clone()
        if task == "matched":
            qa.masked_fill_(batch["matched_labels"] == 0, -100)        # lxmert_pretrain.py:186-188
        elif task == "word_mask":
            qa[0] = -100                                                 # an ignored row outside the matched rule
        labels = dict(word_labels=batch["word_labels"], obj_labels=batch["obj_labels"],
                      matched_labels=batch["matched_labels"], feat_labels=feat_labels, qa_labels=qa)
        model.zero_grad()
        out = model(input_ids=ids, label_dict=labels, task=task, **common)
        out["total_loss"].backward()
        for k, v in out.items():
            rec[f"{k}_{task}"] = v.detach()
        rec[f"qa_labels_{task}"] = qa
        for short, pname in watch.items():
            gr = params[pname].grad
            rec[f"gnorm_{short}_{task}"] = torch.tensor(float("nan") if gr is None else gr.norm().item())
            if gr is not None:
                rec[f"ghead_{short}_{task}"] = gr.flatten()[:16].clone()
    with torch.no_grad():
        feats = synth.visual_feats_from(table, batch["cluster_ids"])
        o = model.bert(input_ids=batch["input_ids"], visual_feats=feats, visual_pos=batch["visual_pos"],
                       attention_mask=batch["attention_mask"], return_dict=True)
        rec["qa_score"] = model.answer_head(o[2])
        rec["pooled"] = o[2]
    save(name, rec)


def golden_sampler(name, d, B, wseed, bseed, n_steps=4):
    """Reference NAR sampling loop (imggen_model.py:199-243) at the BASELINE batch size of config 5 (B = 32), run with
    the reference's ``XLxmertForPretraining`` sub-modules exactly as ``ImggenModel`` chains them; every step's mask,
    max-probability, arg-max, top-2 ids and top-2 logit margin are stored so that the GPU test can replay the masks
    (teacher forcing) and assert the arg-max margin-stratified."""
    model = refshim.build_pretraining_model(
        num_clusters=d.num_clusters, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, task_qa=False)
    sd_bert = P.init_state_dict(P.model_param_specs(d), seed=wseed, randomize_ln_bias=True)
    sd_head = P.init_state_dict(P.objhead_param_specs(d), seed=wseed + 1, randomize_ln_bias=True)
    table = synth.centroid_table(d)
    g = torch.Generator().manual_seed(wseed + 3)
    mask_feat = 0.05 * torch.randn(d.feat_dim, generator=g)
    model.set_visual_embedding(table.clone())
    full = {"bert." + k: v for k, v in sd_bert.items()}
    full.update({"obj_predict_head." + k: v for k, v in sd_head.items() if k != "out_cluster.weight"})
    full["mask_feat"] = mask_feat
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    model.eval()
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    ids, vpos = batch["input_ids"], batch["visual_pos"]
    n_grids = 64
    rec = dict(meta=np.array([B, 20, 64, wseed, bseed, n_steps]),
               weights_checksum=torch.tensor(checksum(sd_bert) + checksum(sd_head), dtype=torch.float64))
    with torch.no_grad():
        for i in range(n_steps):
            n_mask = int((n_steps - i) / n_steps * n_grids)
            if i == 0:
                vis_mask = torch.ones(B, n_grids).long()
                code = torch.zeros(B, n_grids, d.feat_dim)
            else:
                _, lowest_arg = pred_prob.topk(n_mask, dim=1, largest=False)
                vis_mask = torch.zeros(B, n_grids).long()
                vis_mask.scatter_(1, lowest_arg, 1)
            code = torch.where(vis_mask.view(B, n_grids, 1).bool(),
                               model.mask_feat.view(1, 1, -1).to(dtype=code.dtype), code)
            lx = model.bert(input_ids=ids, visual_feats=code, visual_pos=vpos, attention_mask=ids > 0,
                            return_dict=True)
            pl = model.obj_predict_head(lx[1], out_keys=["obj"])["obj"]
            pred_prob, pred_id = torch.softmax(pl, dim=2).max(dim=2)
            code = torch.where(vis_mask.view(B, n_grids, 1).bool(), model.vis_emb(pred_id), code)
            t2v, t2i = pl.topk(2, dim=2)
            rec[f"mask{i}"] = vis_mask.to(torch.uint8)
            rec[f"prob{i}"] = pred_prob
            rec[f"id{i}"] = pred_id.to(torch.int32)
            rec[f"top2_{i}"] = t2i.to(torch.int32)
            rec[f"margin{i}"] = t2v[..., 0] - t2v[..., 1]
        rec["code_sub"] = code[:, ::4, ::128]
    save(name, rec)


def golden_generator(name, B, wseed, bseed):
    """Reference ``Generator`` (eval, noise off) on centroid-table codes."""
    layers = refshim.import_generator_layers()
    G = layers.Generator(base_dim=32, emb_dim=2048, norm_type="spade_in", target_size=256, init_H=8,
                         init_W=8, SN=True, codebook_dim=256)
    sd = P.init_generator_state_dict(seed=wseed)
    G.load_state_dict(sd, strict=True)
    G.eval()
    d = DEFAULT_DIMS
    table = synth.centroid_table(d)
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(table, batch["cluster_ids"])            # [B, 64, 2048]
    emb = code.permute(0, 2, 1).reshape(B, 2048, 8, 8)                      # imggen_model.py:254
    hs = []
    hooks = [rb.register_forward_hook(lambda m, i, o: hs.append(o.detach())) for rb in G.resblocks]
    pre = []
    hooks.append(G.last.register_forward_hook(lambda m, i, o: pre.append(i[0].detach())))
    with torch.no_grad():
        img = G(emb, train=False)
    for h in hooks:
        h.remove()
    rec = dict(meta=np.array([B, wseed, bseed]),
               weights_checksum=torch.tensor(checksum(sd), dtype=torch.float64),
               img_sub=img[:, :, ::4, ::4], pre_tanh_sub=pre[0][:, :, ::4, ::4],
               img_mean=img.mean(), img_absmean=img.abs().mean(),
               saturated_frac=(img.abs() > 0.999).float().mean())
    for i, h in enumerate(hs):
        s = max(1, h.shape[-1] // 16)
        rec[f"h{i}_sub"] = h[:, :, ::s, ::s]
        rec[f"h{i}_absmean"] = h.abs().mean()
    save(name, rec)


def golden_generator_batch(name, B, wseed, bseed, keep=(0, 5, 10, 15)):
    """Reference ``Generator`` at a batch whose 128-pixel GEMM tiles span several images at every stage (B = 16:
    8x8 stage = 64 pixels per image, 2 images per tile).  Stored: image / pre-tanh sub-samples of every image, block
    outputs of the images in ``keep``."""
    layers = refshim.import_generator_layers()
    G = layers.Generator(base_dim=32, emb_dim=2048, norm_type="spade_in", target_size=256, init_H=8,
                         init_W=8, SN=True, codebook_dim=256)
    sd = P.init_generator_state_dict(seed=wseed)
    G.load_state_dict(sd, strict=True)
    G.eval()
    d = DEFAULT_DIMS
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(synth.centroid_table(d), batch["cluster_ids"])
    emb = code.permute(0, 2, 1).reshape(B, 2048, 8, 8)
    hs, pre = [], []
    hooks = [rb.register_forward_hook(lambda m, i, o: hs.append(o.detach())) for rb in G.resblocks]
    hooks.append(G.last.register_forward_hook(lambda m, i, o: pre.append(i[0].detach())))
    with torch.no_grad():
        img = G(emb, train=False)
    for h in hooks:
        h.remove()
    keep = [k for k in keep if k < B]
    rec = dict(meta=np.array([B, wseed, bseed]), keep=np.array(keep),
               weights_checksum=torch.tensor(checksum(sd), dtype=torch.float64),
               img_sub=img[:, :, ::8, ::8], pre_tanh_sub=pre[0][:, :, ::8, ::8],
               img_mean_per_image=img.mean(dim=(1, 2, 3)))
    for i, h in enumerate(hs):
        s_ = max(1, h.shape[-1] // 16)
        rec[f"h{i}_sub"] = h[keep][:, :, ::s_, ::s_]
    save(name, rec)


def golden_generator_noise(name, B, wseed, bseed, noise_seed, noise_weight=0.05):
    """Reference ``Generator.forward(train=True)`` with NON-ZERO noise weights (a trained G; layers.py:56-62) under a
    fixed ``torch.manual_seed``: NoiseInjection draws ``image.new_empty(B,1,R,R).normal_()`` from the global CPU
    generator, noise1 then noise2 of each block in order, so the test re-draws the identical maps from the seed."""
    layers = refshim.import_generator_layers()
    G = layers.Generator(base_dim=32, emb_dim=2048, norm_type="spade_in", target_size=256, init_H=8,
                         init_W=8, SN=True, codebook_dim=256)
    sd = P.init_generator_state_dict(seed=wseed)
    for k in sd:
        if k.endswith("noise1.weight") or k.endswith("noise2.weight"):
            sd[k] = torch.full_like(sd[k], noise_weight)
    G.load_state_dict(sd, strict=True)
    G.eval()
    d = DEFAULT_DIMS
    batch = synth.make_batch(d, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(synth.centroid_table(d), batch["cluster_ids"])
    emb = code.permute(0, 2, 1).reshape(B, 2048, 8, 8)
    hs, pre = [], []
    hooks = [rb.register_forward_hook(lambda m, i, o: hs.append(o.detach())) for rb in G.resblocks]
    hooks.append(G.last.register_forward_hook(lambda m, i, o: pre.append(i[0].detach())))
    torch.manual_seed(noise_seed)
    with torch.no_grad():
        img = G(emb, train=True)
    for h in hooks:
        h.remove()
    rec = dict(meta=np.array([B, wseed, bseed, noise_seed]), noise_weight=np.float32(noise_weight),
               weights_checksum=torch.tensor(checksum(sd), dtype=torch.float64),
               img_sub=img[:, :, ::4, ::4], pre_tanh_sub=pre[0][:, :, ::4, ::4])
    for i, h in enumerate(hs):
        s_ = max(1, h.shape[-1] // 16)
        rec[f"h{i}_sub"] = h[:, :, ::s_, ::s_]
    save(name, rec)


def save(name, rec):
    os.makedirs(OUT, exist_ok=True)
    arrs = {}
    for k, v in rec.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        arrs[k] = v
    import transformers
    arrs["versions"] = np.array([torch.__version__, transformers.__version__])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrs)} arrays")


def main():
    assert refshim.available(), "needs /root/reference"
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    d = DEFAULT_DIMS
    which = sys.argv[1:] or ["model", "ragged", "pretrain", "generator", "sampler", "generator_b16", "generator_noise",
                             "feat_qa"]
    if "feat_qa" in which:
        golden_pretrain_feat_qa("pretrain_feat_qa_b3", d, B=3, wseed=0, bseed=2)
    if "model" in which:
        golden_model("model_b2_l20_v64", d, B=2, L=20, V=64, wseed=0, bseed=0)
    if "ragged" in which:
        golden_model("model_b3_l13_v36", d, B=3, L=13, V=36, wseed=5, bseed=9)
    if "pretrain" in which:
        golden_pretrain("pretrain_b2", d, B=2, wseed=0, bseed=0)
    if "generator" in which:
        golden_generator("generator_b2", B=2, wseed=0, bseed=0)
    if "sampler" in which:
        golden_sampler("sampler_b32", d, B=32, wseed=0, bseed=11)
    if "generator_b16" in which:
        golden_generator_batch("generator_b16", B=16, wseed=0, bseed=3)
    if "generator_noise" in which:
        golden_generator_noise("generator_noise_b2", B=2, wseed=0, bseed=0, noise_seed=1234)


if __name__ == "__main__":
    main()
