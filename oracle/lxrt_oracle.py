"""ORACLE — test infrastructure, not product code.

CPU restatement (plain torch tensor algebra, fp32 or fp64, autograd-capable) of the X-LXMERT hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import this file; the shipped path (``xlxmert_b200``) never does.

Algorithm source.  The arithmetic of this path does **not** live in the reference tree: it is the
third-party dependency ``transformers`` (pinned ``==4.1.1`` in ``/root/reference/requirements.txt:11``;
5.5.0 is what this image has), class ``LxmertModel`` and friends in
``transformers/models/lxmert/modeling_lxmert.py`` (cited below as ``HF:<line>``, 5.5.0 numbering),
called from the reference at ``x-lxmert/src/lxrt/modeling.py:5,80,86,195-206``.  The cluster head and
the pre-training losses are the reference's own (``x-lxmert/src/lxrt/modeling.py:8-53,154-308``), the
sampler step is ``x-lxmert/src/tasks/imggen_model.py:199-243``.

Pinning.  The reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference's own classes run in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` + HF LXMERT, runs them on seeded inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
them, and ``tests/test_oracle_vs_hf.py`` re-checks against the installed HF classes live.

Every function takes ``sd``: a dict ``name -> tensor`` keyed like the reference state dict
*relative to the module it restates* (e.g. ``encoder(sd, …)`` expects ``visn_fc.visn_fc.weight``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def sub(sd: SD, prefix: str) -> SD:
    """View of ``sd`` with ``prefix.`` stripped."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


# ---- primitives ---------------------------------------------------------------------------------

def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """``y = x·Wᵀ + b`` with ``W`` stored ``[out, in]`` (torch.nn.Linear)."""
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-12) -> Tensor:
    """Biased-variance LayerNorm over the last axis, eps 1e-12 (HF:188,281,343,468,472,588)."""
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """erf GeLU — ``ACT2FN["gelu"]`` (HF:331,587)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def extended_mask(attention_mask: Tensor, dtype) -> Tensor:
    """``(1 − mask)·finfo.min`` broadcast to ``[B,1,1,L]`` (HF:766-774)."""
    m = attention_mask[:, None, None, :].to(dtype)
    return (1.0 - m) * torch.finfo(dtype).min


# ---- dropout (training mode) -----------------------------------------------------------------------
# The reference's training forward applies nn.Dropout at HF:189,212 (embeddings), :236,262-266 (attention
# probabilities), :282,286 / :344,348 (dropout(dense(x)) before the residual add) and :474,482 (visual feature
# encoder).  The masks themselves are random; what the oracle pins is the ARITHMETIC around them: a test installs a
# provider that returns, per site, the multiplier tensor (0 or 1/(1−p)) — materialised from the product's own
# counter-based generator (xlx_dropout_mask, site numbering of csrc/dropout.cuh) — and every dropout call below
# multiplies by it, exactly like nn.Dropout does with its own mask.  Without a provider: eval / p = 0.
_DROP_PROVIDER = None


class dropout_masks:
    """``with dropout_masks(provider): ...`` — ``provider(site: int, kind: 'hidden'|'probs', shape) -> Tensor | None``."""

    def __init__(self, provider):
        self.provider = provider

    def __enter__(self):
        global _DROP_PROVIDER
        self.prev, _DROP_PROVIDER = _DROP_PROVIDER, self.provider
        return self

    def __exit__(self, *exc):
        global _DROP_PROVIDER
        _DROP_PROVIDER = self.prev
        return False


def _drop(x: Tensor, site: Optional[int], kind: str) -> Tensor:
    if _DROP_PROVIDER is None or site is None:
        return x
    m = _DROP_PROVIDER(site, kind, tuple(x.shape))
    return x if m is None else x * m.to(x.dtype)


def site_probs(att_block: int, direction: int = 0) -> int:
    return 16 + 4 * att_block + direction


def site_att_out(att_block: int) -> int:
    return 16 + 4 * att_block + 2


def site_ffn_out(n_att: int, ffn_block: int) -> int:
    return 16 + 4 * n_att + ffn_block


# ---- attention blocks ---------------------------------------------------------------------------

def attention(sd: SD, hidden: Tensor, ctx: Tensor, mask: Optional[Tensor], heads: int,
              return_probs: bool = False, drop_site: Optional[int] = None):
    """``LxmertAttention.forward`` (HF:238-274): per-head ``softmax(QKᵀ/√d + mask)·V``."""
    B, Sq, H = hidden.shape
    Sk = ctx.shape[1]
    d = H // heads
    q = linear(hidden, sd["query.weight"], sd["query.bias"]).view(B, Sq, heads, d).transpose(1, 2)
    k = linear(ctx, sd["key.weight"], sd["key.bias"]).view(B, Sk, heads, d).transpose(1, 2)
    v = linear(ctx, sd["value.weight"], sd["value.bias"]).view(B, Sk, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)          # scale after QKᵀ (HF:255-256)
    if mask is not None:
        s = s + mask                                      # additive mask after scaling (HF:258-259)
    p = torch.softmax(s, dim=-1)
    o = (_drop(p, drop_site, "probs") @ v).permute(0, 2, 1, 3).reshape(B, Sq, H)   # dropout(probs)·V, merge heads (HF:262-271)
    return (o, p) if return_probs else o


def attention_output(sd: SD, x: Tensor, residual: Tensor, drop_site: Optional[int] = None,
                     drop_mask: Optional[Tensor] = None) -> Tensor:
    """``LxmertAttentionOutput`` / ``LxmertOutput`` (HF:277-288, 339-350): ``LN(dropout(W·x + b) + residual)``."""
    y = linear(x, sd["dense.weight"], sd["dense.bias"])
    y = y * drop_mask.to(y.dtype) if drop_mask is not None else _drop(y, drop_site, "hidden")
    return layer_norm(y + residual, sd["LayerNorm.weight"], sd["LayerNorm.bias"])


def self_att_layer(sd: SD, x: Tensor, mask, heads: int, blk: Optional[int] = None) -> Tensor:
    """``LxmertSelfAttentionLayer`` (HF:306-324).  ``blk``: attention-block number (dropout sites)."""
    ps, os_ = (None, None) if blk is None else (site_probs(blk), site_att_out(blk))
    return attention_output(sub(sd, "output"), attention(sub(sd, "self"), x, x, mask, heads, drop_site=ps), x, os_)


def cross_att_layer(sd: SD, x: Tensor, ctx: Tensor, ctx_mask, heads: int, probs_site: Optional[int] = None,
                    out_mask: Optional[Tensor] = None) -> Tensor:
    """``LxmertCrossAttentionLayer`` (HF:291-303); the residual is the query-side input."""
    return attention_output(sub(sd, "output"), attention(sub(sd, "att"), x, ctx, ctx_mask, heads, drop_site=probs_site),
                            x, drop_mask=out_mask)


def ffn(sd_inter: SD, sd_out: SD, x: Tensor, drop_site: Optional[int] = None) -> Tensor:
    """``LxmertIntermediate`` + ``LxmertOutput`` (HF:327-350)."""
    h = gelu_erf(linear(x, sd_inter["dense.weight"], sd_inter["dense.bias"]))
    return attention_output(sd_out, h, x, drop_site)


def layer(sd: SD, x: Tensor, mask, heads: int, blk: Optional[int] = None, n_att: int = 0) -> Tensor:
    """``LxmertLayer`` (HF:353-366): self-attention block then FFN block (``blk`` = its number in both plans)."""
    a = self_att_layer(sub(sd, "attention"), x, mask, heads, blk)
    return ffn(sub(sd, "intermediate"), sub(sd, "output"), a, None if blk is None else site_ffn_out(n_att, blk))


def xlayer(sd: SD, lang: Tensor, lmask, vis: Tensor, vmask, heads: int, att_blk: Optional[int] = None,
           ffn_blk: Optional[int] = None, n_att: int = 0):
    """``LxmertXLayer`` (HF:369-457): cross (shared weights, both directions read the layer
    inputs, HF:385-406) → self (HF:408-412) → FFN (HF:414-423).  ``att_blk`` / ``ffn_blk``: number of this layer's
    first attention / FFN block (cross, lang-self, vis-self / lang, vis) for the dropout sites; the product runs the
    shared cross-attention output layer once over [language rows | vision rows], so ONE site covers both."""
    xa = sub(sd, "visual_attention")
    lm = vm = None
    ps0 = ps1 = None
    if att_blk is not None and _DROP_PROVIDER is not None:
        B, L, H = lang.shape
        V = vis.shape[1]
        joint = _DROP_PROVIDER(site_att_out(att_blk), "hidden", (B * (L + V), H))
        if joint is not None:
            lm, vm = joint[:B * L].reshape(B, L, H), joint[B * L:].reshape(B, V, H)
        ps0, ps1 = site_probs(att_blk, 0), site_probs(att_blk, 1)
    l1 = cross_att_layer(xa, lang, vis, vmask, heads, ps0, lm)
    v1 = cross_att_layer(xa, vis, lang, lmask, heads, ps1, vm)
    sl = None if att_blk is None else att_blk + 1
    sv = None if att_blk is None else att_blk + 2
    l2 = self_att_layer(sub(sd, "lang_self_att"), l1, lmask, heads, sl)
    v2 = self_att_layer(sub(sd, "visn_self_att"), v1, vmask, heads, sv)
    l3 = ffn(sub(sd, "lang_inter"), sub(sd, "lang_output"), l2, None if ffn_blk is None else site_ffn_out(n_att, ffn_blk))
    v3 = ffn(sub(sd, "visn_inter"), sub(sd, "visn_output"), v2,
             None if ffn_blk is None else site_ffn_out(n_att, ffn_blk + 1))
    return l3, v3


def visual_feature_encoder(sd: SD, feats: Tensor, pos: Tensor) -> Tensor:
    """``LxmertVisualFeatureEncoder`` (HF:476-484): ``dropout((LN(Wf·x) + LN(Wp·pos)) / 2)``."""
    x = layer_norm(linear(feats, sd["visn_fc.weight"], sd["visn_fc.bias"]),
                   sd["visn_layer_norm.weight"], sd["visn_layer_norm.bias"])
    y = layer_norm(linear(pos, sd["box_fc.weight"], sd["box_fc.bias"]),
                   sd["box_layer_norm.weight"], sd["box_layer_norm.bias"])
    return _drop((x + y) / 2, 1, "hidden")


def encoder(sd: SD, lang: Tensor, lmask, feats: Tensor, pos: Tensor, vmask=None, *, heads: int,
            n_l: int, n_r: int, n_x: int):
    """``LxmertEncoder.forward`` (HF:506-565).  Returns (lang_states, vis_states) lists of every
    layer's output: 9 L + 5 X language states, 5 R + 5 X vision states."""
    vis = visual_feature_encoder(sub(sd, "visn_fc"), feats, pos)
    lang_states, vis_states = [], []
    n_att = n_l + n_r + 3 * n_x           # block numbering of the dropout sites (csrc/dropout.cuh)
    for i in range(n_l):
        lang = layer(sub(sd, f"layer.{i}"), lang, lmask, heads, i, n_att)
        lang_states.append(lang)
    for i in range(n_r):
        vis = layer(sub(sd, f"r_layers.{i}"), vis, vmask, heads, n_l + i, n_att)
        vis_states.append(vis)
    for i in range(n_x):
        lang, vis = xlayer(sub(sd, f"x_layers.{i}"), lang, lmask, vis, vmask, heads, n_l + n_r + 3 * i,
                           n_l + n_r + 2 * i, n_att)
        lang_states.append(lang)
        vis_states.append(vis)
    return lang_states, vis_states


def embeddings(sd: SD, input_ids: Tensor, token_type_ids: Optional[Tensor] = None) -> Tensor:
    """``LxmertEmbeddings`` (HF:191-214); dropout (HF:212) only under ``dropout_masks``."""
    B, L = input_ids.shape
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    pos = torch.arange(L, device=input_ids.device).unsqueeze(0).expand(B, L)
    # all three tables are nn.Embedding(padding_idx=0) (HF:184-186): row 0 is looked up like any other
    # row in forward but receives no gradient — that includes position 0 and token type 0.
    emb = torch.nn.functional.embedding
    e = (emb(input_ids, sd["word_embeddings.weight"], padding_idx=0)
         + emb(pos, sd["position_embeddings.weight"], padding_idx=0)
         + emb(token_type_ids, sd["token_type_embeddings.weight"], padding_idx=0))
    return _drop(layer_norm(e, sd["LayerNorm.weight"], sd["LayerNorm.bias"]), 0, "hidden")


def pooler(sd: SD, lang: Tensor) -> Tensor:
    """``LxmertPooler`` (HF:574-580)."""
    return torch.tanh(linear(lang[:, 0], sd["dense.weight"], sd["dense.bias"]))


def lxmert_model(sd: SD, input_ids: Tensor, visual_feats: Tensor, visual_pos: Tensor,
                 attention_mask: Optional[Tensor] = None, token_type_ids=None, *, heads: int = 12,
                 n_l: int = 9, n_r: int = 5, n_x: int = 5):
    """``LxmertModel.forward`` (HF:699-830) with ``visual_attention_mask=None`` (always the case in
    the reference).  Returns ``(lang_out, vis_out, pooled, lang_states, vis_states)``."""
    dtype = sd["pooler.dense.weight"].dtype
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    lmask = extended_mask(attention_mask, dtype)
    emb = embeddings(sub(sd, "embeddings"), input_ids, token_type_ids)
    ls, vs = encoder(sub(sd, "encoder"), emb, lmask, visual_feats, visual_pos, None, heads=heads,
                     n_l=n_l, n_r=n_r, n_x=n_x)
    return ls[-1], vs[-1], pooler(sub(sd, "pooler"), ls[-1]), ls, vs


# ---- heads, losses, sampler step ---------------------------------------------------------------

def head_transform(sd: SD, h: Tensor) -> Tensor:
    """``LxmertPredictionHeadTransform`` (HF:583-594): ``LN(gelu(W·h + b))``."""
    return layer_norm(gelu_erf(linear(h, sd["dense.weight"], sd["dense.bias"])),
                      sd["LayerNorm.weight"], sd["LayerNorm.bias"])


def obj_head(sd: SD, vis_out: Tensor):
    """``lxrt.modeling.LxmertVisualObjHead.forward`` (x-lxmert/src/lxrt/modeling.py:38-53), cluster
    mode: returns ``(feat [B,V,F], obj_logits [B,V,C])``."""
    t = head_transform(sub(sd, "transform"), vis_out)
    feat = linear(t, sd["linear_feat.weight"], sd["linear_feat.bias"])
    return feat, linear(feat, sd["out_cluster.weight"], sd["out_cluster.bias"])


def cross_entropy_mean(logits: Tensor, labels: Tensor, ignore_index: int = -100) -> Tensor:
    """``CrossEntropyLoss()`` (modeling.py:99,102): mean over rows whose label ≠ −100."""
    lse = torch.logsumexp(logits, dim=-1)
    keep = labels != ignore_index
    safe = labels.clamp(min=0)
    nll = lse - logits.gather(-1, safe.unsqueeze(-1)).squeeze(-1)
    return (nll * keep).sum() / keep.sum()


def obj_loss(sd_head: SD, vis_out: Tensor, obj_labels: Tensor) -> Tensor:
    """Masked-cell cluster prediction loss (modeling.py:244-258)."""
    _, logits = obj_head(sd_head, vis_out)
    return cross_entropy_mean(logits.reshape(-1, logits.shape[-1]), obj_labels.reshape(-1))


def feat_loss(pred_feat: Tensor, feat_labels: Tensor, vis_mask: Tensor) -> Tensor:
    """SmoothL1(β=1) → mean over 2048 → masked mean per sample → batch mean (modeling.py:270-284)."""
    d = (pred_feat - feat_labels).abs()
    h = torch.where(d < 1.0, 0.5 * d * d, d - 0.5).mean(dim=2)
    m = vis_mask.to(h.dtype)
    return ((h * m).sum(1) / m.sum(1).clamp(min=1)).mean()


def answer_head(sd: SD, pooled: Tensor) -> Tensor:
    """HF ``LxmertVisualAnswerHead`` (HF modeling_lxmert.py:610-623), the reference's ``answer_head``
    (x-lxmert/src/lxrt/modeling.py:90,289): ``Linear(H, 2H) → GeLU → LayerNorm(2H, eps 1e-12) → Linear(2H, labels)``
    on the pooled output.  State-dict keys: ``logit_fc.{0,2,3}.{weight,bias}``."""
    t = layer_norm(gelu_erf(linear(pooled, sd["logit_fc.0.weight"], sd["logit_fc.0.bias"])),
                   sd["logit_fc.2.weight"], sd["logit_fc.2.bias"])
    return linear(t, sd["logit_fc.3.weight"], sd["logit_fc.3.bias"])


def qa_loss(sd: SD, pooled: Tensor, qa_labels: Tensor):
    """``CrossEntropyLoss()(answer_score, ans)`` and ``answer_score.max(1)`` ids (modeling.py:287-299)."""
    score = answer_head(sd, pooled)
    return cross_entropy_mean(score, qa_labels.reshape(-1)), score.argmax(dim=1)


def lm_head(sd_cls: SD, lang_out: Tensor, pooled: Tensor):
    """``LxmertPreTrainingHeads`` (HF:656-665, 597-607); decoder weight tied to word embeddings."""
    t = head_transform(sub(sd_cls, "predictions.transform"), lang_out)
    scores = linear(t, sd_cls["predictions.decoder.weight"], None) + sd_cls["predictions.bias"]
    rel = linear(pooled, sd_cls["seq_relationship.weight"], sd_cls["seq_relationship.bias"])
    return scores, rel


def mask_visual_feats(feats: Tensor, vis_mask: Tensor, mask_feat: Tensor) -> Tensor:
    """``torch.where(vis_mask, mask_feat, feats)`` on the 2048-d input (modeling.py:190-193)."""
    return torch.where(vis_mask.bool().unsqueeze(-1), mask_feat.view(1, 1, -1).to(feats.dtype), feats)


def sampler_predict(logits: Tensor):
    """``softmax(logits, 2).max(2)`` → (pred_prob, pred_id); first index wins ties
    (x-lxmert/src/tasks/imggen_model.py:232-235)."""
    return torch.softmax(logits, dim=2).max(dim=2)


def nar_n_mask(step: int, n_steps: int, n_grids: int = 64) -> int:
    """Mask-predict linear decay (imggen_model.py:200-202)."""
    return int((n_steps - step) / n_steps * n_grids)


def sampler_step(sd_bert: SD, sd_head: SD, table: Tensor, mask_feat: Tensor, input_ids: Tensor,
                 visual_pos: Tensor, code: Tensor, vis_mask: Tensor, **kw):
    """One iteration of ``sample_image_NAR`` with a *given* ``vis_mask`` (teacher-forced, SURVEY
    App. A): imggen_model.py:215-243.  Returns (new_code, pred_prob, pred_id)."""
    code_in = mask_visual_feats(code, vis_mask, mask_feat)
    _, vis_out, _, _, _ = lxmert_model(sd_bert, input_ids, code_in, visual_pos, input_ids > 0, **kw)
    _, logits = obj_head(sd_head, vis_out)
    pred_prob, pred_id = sampler_predict(logits)
    new_code = torch.where(vis_mask.bool().unsqueeze(-1), table[pred_id], code_in)
    return new_code, pred_prob, pred_id


def to_dtype(sd: SD, dtype) -> SD:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
