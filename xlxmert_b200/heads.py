"""Pre-training heads on the sm_100a library, under the reference's module and parameter names.

* ``B200LxmertVisualObjHead`` — ``lxrt.modeling.LxmertVisualObjHead`` (``x-lxmert/src/lxrt/modeling.py:8-53``,
  cluster mode): ``forward(hidden_states, out_keys=[])`` → ``{'feat': …, 'obj': …}``; plus the two fused entry
  points the callers actually need: ``loss(hidden, obj_labels)`` (head + ``CrossEntropyLoss``,
  ``modeling.py:244-258``) and ``predict(hidden)`` (head + ``softmax(2).max(2)``, ``tasks/imggen_model.py:228-235``).
* ``B200LxmertPreTrainingHeads`` — HF ``LxmertPreTrainingHeads`` (HF ``modeling_lxmert.py:656-665``) with the
  4.1.1 constructor the reference uses (``modeling.py:86``: decoder weight tied to the word embeddings).
* ``B200LxmertVisualAnswerHead`` — HF ``LxmertVisualAnswerHead`` (HF ``modeling_lxmert.py:610-623``): the
  reference's ``answer_head`` under ``--taskQA`` (``modeling.py:89-90,286-299``) and the fine-tune models' answer
  classifier (``tasks/vqa_model.py:16-19``).

``forward`` of every head is differentiable (the backward takes the upstream logits / feature gradient); the fused
``*loss`` entry points additionally keep the soft-max + cross-entropy (and SmoothL1) inside the library.

No PyTorch fallback: every method raises if the CUDA library is missing or the tensors are not on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .config import LxmertDims


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _parr(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Transform(nn.Module):      # LxmertPredictionHeadTransform (HF:583-594)
    def __init__(self, H):
        super().__init__()
        self.dense = nn.Linear(H, H)
        self.LayerNorm = nn.LayerNorm(H, eps=1e-12)


class _FusedHead:
    """Shared machinery of the two big heads: prepared-weight cache, workspace, C-ABI calls."""

    def __init__(self, kind: str, dims: LxmertDims, classes: int, passes: int):
        self.kind, self.dims, self.classes, self.passes = kind, dims, classes, passes
        self.cdims = _lib.XlxDims.from_dims(dims)
        self._prep = None
        self._prep_key = None

    def fn(self, name):
        return getattr(_lib.load(), f"xlx_{self.kind}_{name}")

    def prepared(self, params, force: bool = False):
        """Split-bf16 weight copies; ``force`` (training forwards) rebuilds them unconditionally — see
        ``B200LxmertEncoder._prepared`` for why a version-counter key cannot be trusted across optimiser steps."""
        key = None if force else tuple((p.data_ptr(), p._version) for p in params)
        dev = params[0].device
        if self._prep is None or self._prep.device != dev:
            n = self.fn("prep_bytes")(C.byref(self.cdims), self.classes)
            if n == 0:
                raise _lib.XlxError(f"xlx_{self.kind}_prep_bytes", -20)
            self._prep = torch.empty(n, dtype=torch.uint8, device=dev)
            self._prep_key = None
        if force or key != self._prep_key:
            for p in params:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("head parameters must be contiguous fp32")
            rc = self.fn("prepare")(C.byref(self.cdims), self.classes, _parr(params), self._prep.data_ptr(), _stream())
            _lib.check(f"xlx_{self.kind}_prepare", rc)
            self._prep_key = key
        return self._prep

    def workspace(self, M, dev):
        n = self.fn("workspace_bytes")(C.byref(self.cdims), self.classes, M)
        return torch.empty(n, dtype=torch.uint8, device=dev), n

    def forward(self, params, hidden, labels=None, feat=None, want_feat=False, want_logits=False, want_pred=False,
                training=False):
        """→ dict(feat, logits, loss, feat_loss, pred_prob, pred_id, ws, nws).  ``feat`` = ``(target [M,F], weight [M])``
        of the feature-regression loss (cluster head only).  ``training``: the workspace must survive until a backward
        (implied by ``labels`` / ``feat``)."""
        if not hidden.is_cuda:
            raise RuntimeError("the prediction heads run on CUDA (sm_100a) only; there is no CPU fallback")
        d = self.dims
        lead = hidden.shape[:-1]
        h2 = hidden.reshape(-1, d.hidden).contiguous().float()
        M, dev = h2.shape[0], h2.device
        training = training or labels is not None or feat is not None
        prep = self.prepared(params, force=training)      # a differentiated call is a training step: weights just changed
        if not training:        # inference (sampler loop): one grow-only workspace, stream-ordered reuse
            nws = self.fn("workspace_bytes")(C.byref(self.cdims), self.classes, M)
            ws = getattr(self, "_ws_infer", None)
            if ws is None or ws.numel() < nws or ws.device != dev:
                ws = self._ws_infer = torch.empty(nws, dtype=torch.uint8, device=dev)
        else:                   # training: the workspace carries the saved activations to the backward
            ws, nws = self.workspace(M, dev)
        out = dict(ws=ws, nws=nws, prep=prep, M=M, feat=None, logits=None, loss=None, feat_loss=None, pred_prob=None,
                   pred_id=None, labels=None, feat_in=feat)
        if labels is not None:
            labels = labels.reshape(-1).contiguous()
            out["loss"] = torch.empty((), device=dev, dtype=torch.float32)
            out["labels"] = labels
        if want_logits:
            out["logits"] = torch.empty(*lead, self.classes, device=dev, dtype=torch.float32)
        if want_pred:
            out["pred_prob"] = torch.empty(*lead, device=dev, dtype=torch.float32)
            out["pred_id"] = torch.empty(*lead, device=dev, dtype=torch.int64)
        if self.kind == "objhead":
            if want_feat:
                out["feat"] = torch.empty(*lead, d.feat_dim, device=dev, dtype=torch.float32)
            if feat is not None:
                out["feat_loss"] = torch.empty((), device=dev, dtype=torch.float32)
            ft, fw = feat if feat is not None else (None, None)
            rc = self.fn("fwd")(C.byref(self.cdims), self.classes, _parr(params), prep.data_ptr(), M, h2.data_ptr(),
                                _ptr(labels), _ptr(ft), _ptr(fw), _ptr(out["feat"]), _ptr(out["logits"]),
                                _ptr(out["loss"]), _ptr(out["feat_loss"]), _ptr(out["pred_prob"]), _ptr(out["pred_id"]),
                                ws.data_ptr(), nws, self.passes, _stream())
        else:
            if feat is not None or want_feat:
                raise ValueError("only the cluster head has a feature layer")
            rc = self.fn("fwd")(C.byref(self.cdims), self.classes, _parr(params), prep.data_ptr(), M, h2.data_ptr(),
                                _ptr(labels), _ptr(out["logits"]), _ptr(out["loss"]), _ptr(out["pred_prob"]),
                                _ptr(out["pred_id"]), ws.data_ptr(), nws, self.passes, _stream())
        _lib.check(f"xlx_{self.kind}_fwd", rc)
        return out

    def backward(self, params, saved, hidden_shape, d_loss=None, d_logits=None, d_feat_loss=None, d_feat=None):
        """→ (d_hidden, grads).  Logits gradient: the fused cross-entropy's (``saved['labels']`` + ``d_loss``) or
        ``d_logits``; cluster head: + the regression loss (``saved['feat_in']`` + ``d_feat_loss``) or ``d_feat``."""
        dev = saved["ws"].device
        labels = saved["labels"]
        f32 = lambda t: None if t is None else t.contiguous().float()
        d_loss, d_logits, d_feat_loss, d_feat = f32(d_loss), f32(d_logits), f32(d_feat_loss), f32(d_feat)
        if labels is None:
            d_loss = None
        d_hidden = torch.empty(hidden_shape, device=dev, dtype=torch.float32)
        have_dlogits = labels is not None or d_logits is not None
        skip = set()
        if self.kind == "objhead":
            skip = {6} if have_dlogits else {6, 7}       # frozen centroid table; no classifier gradient at all
        grads = [None if i in skip else torch.empty_like(p) for i, p in enumerate(params)]
        if self.kind == "objhead":
            ft, fw = saved["feat_in"] if saved["feat_in"] is not None else (None, None)
            if ft is None:
                d_feat_loss = None
            rc = self.fn("bwd")(C.byref(self.cdims), self.classes, _parr(params), saved["prep"].data_ptr(), saved["M"],
                                _ptr(labels), _ptr(d_loss), _ptr(d_logits), _ptr(ft), _ptr(fw), _ptr(d_feat_loss),
                                _ptr(d_feat), d_hidden.data_ptr(), _parr(grads), saved["ws"].data_ptr(), saved["nws"],
                                self.passes, _stream())
        else:
            rc = self.fn("bwd")(C.byref(self.cdims), self.classes, _parr(params), saved["prep"].data_ptr(), saved["M"],
                                _ptr(labels), _ptr(d_loss), _ptr(d_logits), d_hidden.data_ptr(), _parr(grads),
                                saved["ws"].data_ptr(), saved["nws"], self.passes, _stream())
        _lib.check(f"xlx_{self.kind}_bwd", rc)
        return d_hidden, grads


_KEEP = ("ws", "nws", "prep", "M", "labels", "feat_in")     # what a backward needs; outputs stay off ctx (see below)


def _param_grads(grads, params):
    return [g if (g is not None and p.requires_grad) else None for g, p in zip(grads, params)]


class _HeadLossFn(torch.autograd.Function):
    """(hidden, *params) → ``(loss, feat_loss, pred_id)``: ``CrossEntropyLoss(head(hidden), labels)`` as a 0-d tensor
    (``None`` without labels), the cluster head's feature-regression loss (``None`` without ``feat``) and the arg-max
    ids (``None`` unless asked for)."""

    @staticmethod
    def forward(ctx, fused: _FusedHead, labels, feat, want_pred, hidden, *params):
        out = fused.forward(list(params), hidden, labels=labels, feat=feat, want_pred=want_pred)
        if any(ctx.needs_input_grad):
            # everything but the returned tensors: an output stored on ctx would tie the workspace (GBs of logits) into
            # a reference cycle that only the cyclic garbage collector frees
            ctx.fused, ctx.saved, ctx.params, ctx.hshape = fused, {k: out[k] for k in _KEEP}, params, hidden.shape
        if out["pred_id"] is not None:
            ctx.mark_non_differentiable(out["pred_id"])
        return out["loss"], out["feat_loss"], out["pred_id"]

    @staticmethod
    def backward(ctx, d_loss, d_feat_loss, _d_pred):
        d_hidden, grads = ctx.fused.backward(list(ctx.params), ctx.saved, ctx.hshape, d_loss=d_loss,
                                             d_feat_loss=d_feat_loss)
        ctx.saved = None
        return (None, None, None, None, d_hidden, *_param_grads(grads, ctx.params))


class _HeadOutputsFn(torch.autograd.Function):
    """(hidden, *params) → ``(feat, logits)`` (either may be ``None``), differentiable: the reference's plain
    ``head(hidden)`` call for callers that bring their own loss (``tasks/vqa_model.py`` + BCE, ``out_keys=['feat']``)."""

    @staticmethod
    def forward(ctx, fused: _FusedHead, want_feat, want_logits, hidden, *params):
        out = fused.forward(list(params), hidden, want_feat=want_feat, want_logits=want_logits, training=True)
        ctx.set_materialize_grads(False)
        ctx.fused, ctx.saved, ctx.params, ctx.hshape = fused, {k: out[k] for k in _KEEP}, params, hidden.shape
        return out["feat"], out["logits"]

    @staticmethod
    def backward(ctx, d_feat, d_logits):
        if d_feat is None and d_logits is None:
            return (None,) * (4 + len(ctx.params))
        C_ = ctx.fused.classes
        if d_logits is not None:
            d_logits = d_logits.reshape(-1, C_)
        if d_feat is not None:
            d_feat = d_feat.reshape(-1, d_feat.shape[-1])
        d_hidden, grads = ctx.fused.backward(list(ctx.params), ctx.saved, ctx.hshape, d_logits=d_logits, d_feat=d_feat)
        ctx.saved = None
        return (None, None, None, d_hidden, *_param_grads(grads, ctx.params))


def _differentiable(hidden, params) -> bool:
    return torch.is_grad_enabled() and (hidden.requires_grad or any(p.requires_grad for p in params))


class _LabelledRowsFn(torch.autograd.Function):
    """hidden [M, H] → the rows whose label is not −100 (ascending); backward scatters into zeros."""

    @staticmethod
    def forward(ctx, hidden, rows, n):
        M, H = hidden.shape
        out = torch.empty(n, H, device=hidden.device, dtype=torch.float32)
        rc = _lib.load().xlx_gather_rows(hidden.data_ptr(), None, rows.data_ptr(), n, H, out.data_ptr(), None, _stream())
        _lib.check("xlx_gather_rows", rc)
        ctx.rows, ctx.n, ctx.M = rows, n, M
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        d_hidden = torch.empty(ctx.M, d_out.shape[1], device=d_out.device, dtype=torch.float32)
        rc = _lib.load().xlx_scatter_rows(d_out.data_ptr(), ctx.rows.data_ptr(), ctx.n, ctx.M, d_out.shape[1],
                                          d_hidden.data_ptr(), _stream())
        _lib.check("xlx_scatter_rows", rc)
        return d_hidden, None, None


def compact_rows(hidden, selector):
    """Rows of ``hidden`` [.., H] whose ``selector`` entry (int64, flattened) is not −100.  → ``(hidden', rows, n)``:
    ``rows`` holds the ascending indices (first ``n`` valid) or is ``None`` when nothing was dropped (every row
    selected, or none: the all-ignored loss is NaN like the reference's).  Costs one 4-byte device→host read (the row
    count sizes the head's GEMMs)."""
    if not hidden.is_cuda:
        raise RuntimeError("the prediction heads run on CUDA (sm_100a) only; there is no CPU fallback")
    H = hidden.shape[-1]
    flat = selector.reshape(-1).contiguous()
    M = flat.numel()
    rows = torch.empty(M, dtype=torch.int64, device=hidden.device)
    count = torch.empty(1, dtype=torch.int32, device=hidden.device)
    lib = _lib.load()
    _lib.check("xlx_labelled_rows", lib.xlx_labelled_rows(flat.data_ptr(), M, -100, rows.data_ptr(), count.data_ptr(),
                                                          _stream()))
    n = int(count.item())
    if n == 0 or n == M:
        return hidden, None, M
    h2 = hidden.reshape(M, H)
    if h2.dtype != torch.float32 or not h2.is_contiguous():
        h2 = h2.contiguous().float()
    return _LabelledRowsFn.apply(h2, rows, n), rows, n


def pick_labels(labels, rows, n):
    flat = labels.reshape(-1).contiguous()
    if rows is None:
        return flat
    picked = torch.empty(n, dtype=torch.int64, device=flat.device)
    _lib.check("xlx_gather_rows", _lib.load().xlx_gather_rows(None, flat.data_ptr(), rows.data_ptr(), n, 0, None,
                                                              picked.data_ptr(), _stream()))
    return picked


def pick_rows(x, rows, n):
    """``x [M, cols]`` fp32 → its ``rows`` (no gradient: targets)."""
    x = x.reshape(-1, x.shape[-1])
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.contiguous().float()
    if rows is None:
        return x
    out = torch.empty(n, x.shape[1], device=x.device, dtype=torch.float32)
    _lib.check("xlx_gather_rows", _lib.load().xlx_gather_rows(x.data_ptr(), None, rows.data_ptr(), n, x.shape[1],
                                                              out.data_ptr(), None, _stream()))
    return out


def labelled_rows(hidden, labels):
    """Drop the rows CrossEntropyLoss ignores (label −100, modeling.py:99,253-256) before a masked-prediction head:
    same loss, same gradients, none of the head's GEMM work for rows that cannot contribute.  Returns
    ``(hidden, labels)`` unchanged when every row is labelled or none is."""
    h2, rows, n = compact_rows(hidden, labels)
    if rows is None:
        return hidden, labels
    return h2, pick_labels(labels, rows, n)


class B200LxmertVisualObjHead(nn.Module):
    def __init__(self, dims: LxmertDims, num_clusters: Optional[int] = None, passes: int = 3,
                 source: Optional[nn.Module] = None):
        super().__init__()
        C_ = num_clusters if num_clusters is not None else dims.num_clusters
        if source is not None:
            self.transform, self.linear_feat, self.out_cluster = source.transform, source.linear_feat, source.out_cluster
            C_ = source.out_cluster.out_features
        else:
            self.transform = _Transform(dims.hidden)
            self.linear_feat = nn.Linear(dims.hidden, dims.feat_dim)
            self.out_cluster = nn.Linear(dims.feat_dim, C_)
        self.cluster_out = True
        self.compact_rows = True      # run the head on labelled rows only (see labelled_rows)
        # which outputs `forward` returns by default (modeling.py:12-25,41-50).  --visualLosses obj (pretrain.bash);
        # add "feat" for the published default `--visualLosses obj,feat` (param.py:123)
        self.visual_losses = {"obj": {"shape": (-1,), "num": C_}}
        self._fused = _FusedHead("objhead", dims, C_, passes)
        self._feat_dim = dims.feat_dim

    def _params(self) -> List[torch.Tensor]:
        t = self.transform
        return [t.dense.weight, t.dense.bias, t.LayerNorm.weight, t.LayerNorm.bias, self.linear_feat.weight,
                self.linear_feat.bias, self.out_cluster.weight, self.out_cluster.bias]

    def forward(self, hidden_states, out_keys=[]):
        """Reference signature (modeling.py:38-53) → ``{'feat': …, 'obj': …}`` for the keys in ``visual_losses`` and
        ``out_keys``; differentiable when gradients are enabled.  Training loops should prefer :meth:`losses`, which
        keeps soft-max + cross-entropy inside the library and never materialises a logits gradient in PyTorch."""
        keys = set(self.visual_losses) | set(out_keys)
        params = self._params()
        if _differentiable(hidden_states, params):
            feat, logits = _HeadOutputsFn.apply(self._fused, "feat" in keys, "obj" in keys, hidden_states, *params)
        else:
            out = self._fused.forward(params, hidden_states, want_feat="feat" in keys, want_logits="obj" in keys)
            feat, logits = out["feat"], out["logits"]
        res = {}
        if "feat" in keys:
            res["feat"] = feat
        if "obj" in keys:
            res["obj"] = logits
        return res

    def losses(self, hidden_states, obj_labels=None, feat_labels=None, vis_mask=None):
        """The visual losses of the ``vis_mask`` task (modeling.py:237-284) → ``{'obj': …, 'feat': …}`` (0-d,
        differentiable; a key is present iff its labels were given):

        * ``obj``:  ``CrossEntropyLoss()(obj_logit.view(B·V, C), obj_labels.flatten())`` (:244-258);
        * ``feat``: ``SmoothL1Loss(reduction='none')(pred_feat, feat_labels).mean(2)``, masked by ``vis_mask``, divided
          by ``vis_mask.sum(1).clamp(min=1)``, batch mean (:270-284).
        """
        if obj_labels is None and feat_labels is None:
            raise ValueError("losses() needs obj_labels and / or feat_labels")
        B, V = hidden_states.shape[:2]
        dev = hidden_states.device
        mask = None
        if feat_labels is not None:
            if vis_mask is None:
                raise ValueError("the feature-regression loss needs vis_mask (modeling.py:279)")
            mask = vis_mask.reshape(B, V).to(torch.bool).contiguous()
        rows, n, hidden = None, B * V, hidden_states
        if self.compact_rows:
            # rows that can contribute: labelled for the cross-entropy, masked for the regression
            if mask is None:
                selector = obj_labels
            else:
                keep = mask if obj_labels is None else (mask | (obj_labels.reshape(B, V) != -100))
                selector = torch.where(keep, 0, -100)
            hidden, rows, n = compact_rows(hidden_states, selector)
        labels = None if obj_labels is None else pick_labels(obj_labels, rows, n)
        feat = None
        if feat_labels is not None:
            target = pick_rows(feat_labels.reshape(B * V, -1), rows, n)
            weight = torch.empty(n, device=dev, dtype=torch.float32)
            rc = _lib.load().xlx_feat_row_weight(mask.data_ptr(), B, V, self._feat_dim, _ptr(rows), n,
                                                 weight.data_ptr(), _stream())
            _lib.check("xlx_feat_row_weight", rc)
            feat = (target, weight)
        obj, fl, _ = _HeadLossFn.apply(self._fused, labels, feat, False, hidden, *self._params())
        res = {}
        if obj_labels is not None:
            res["obj"] = obj
        if feat_labels is not None:
            res["feat"] = fl
        return res

    def loss(self, hidden_states, obj_labels):
        """``CrossEntropyLoss()(obj_logit.view(B·V, C), obj_label.flatten())`` (modeling.py:253-256), differentiable."""
        return self.losses(hidden_states, obj_labels=obj_labels)["obj"]

    @torch.no_grad()
    def predict(self, hidden_states):
        """``softmax(obj_logits, 2).max(2)`` → ``(pred_prob, pred_id)`` (imggen_model.py:228-235)."""
        out = self._fused.forward(self._params(), hidden_states, want_pred=True)
        return out["pred_prob"], out["pred_id"]


class _MatchLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, labels, pooled, W, b):
        lib = _lib.load()
        B = pooled.shape[0]
        dev = pooled.device
        pooled = pooled.contiguous().float()
        labels = labels.reshape(-1).contiguous()
        scores = torch.empty(B, 2, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        scratch = torch.empty(lib.xlx_matchhead_scratch_floats(B), device=dev, dtype=torch.float32)
        rc = lib.xlx_matchhead_fwd(C.byref(mod._cdims), B, pooled.data_ptr(), W.data_ptr(), b.data_ptr(),
                                   labels.data_ptr(), scores.data_ptr(), loss.data_ptr(), scratch.data_ptr(), _stream())
        _lib.check("xlx_matchhead_fwd", rc)
        ctx.mod, ctx.saved = mod, (pooled, W, labels, scores, scratch)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        lib = _lib.load()
        pooled, W, labels, scores, scratch = ctx.saved
        B = pooled.shape[0]
        d_loss = d_loss.contiguous().float()
        d_pooled, dW = torch.empty_like(pooled), torch.empty_like(W)
        db = torch.empty(2, device=W.device, dtype=torch.float32)
        rc = lib.xlx_matchhead_bwd(C.byref(ctx.mod._cdims), B, pooled.data_ptr(), W.data_ptr(), labels.data_ptr(),
                                   scores.data_ptr(), d_loss.data_ptr(), d_pooled.data_ptr(), dW.data_ptr(),
                                   db.data_ptr(), scratch.data_ptr(), _stream())
        _lib.check("xlx_matchhead_bwd", rc)
        return None, None, d_pooled, dW, db


def _match_scores(mod, pooled_output):
    lib = _lib.load()
    B = pooled_output.shape[0]
    pooled = pooled_output.contiguous().float()
    rel = torch.empty(B, 2, device=pooled.device, dtype=torch.float32)
    rc = lib.xlx_matchhead_fwd(C.byref(mod._cdims), B, pooled.data_ptr(), mod.seq_relationship.weight.data_ptr(),
                               mod.seq_relationship.bias.data_ptr(), None, rel.data_ptr(), None, None, _stream())
    _lib.check("xlx_matchhead_fwd", rc)
    return rel


class _MatchScoresFn(torch.autograd.Function):
    """pooled → ``seq_relationship(pooled)`` [B, 2], differentiable."""

    @staticmethod
    def forward(ctx, mod, pooled, W, b):
        ctx.mod = mod
        ctx.save_for_backward(pooled, W)
        return _match_scores(mod, pooled)

    @staticmethod
    def backward(ctx, d_rel):
        pooled, W = ctx.saved_tensors
        pooled = pooled.contiguous().float()
        d_rel = d_rel.contiguous().float()
        B = pooled.shape[0]
        d_pooled, dW = torch.empty_like(pooled), torch.empty_like(W)
        db = torch.empty(2, device=W.device, dtype=torch.float32)
        rc = _lib.load().xlx_matchhead_bwd_scores(C.byref(ctx.mod._cdims), B, pooled.data_ptr(), W.data_ptr(),
                                                  d_rel.data_ptr(), d_pooled.data_ptr(), dW.data_ptr(), db.data_ptr(),
                                                  _stream())
        _lib.check("xlx_matchhead_bwd_scores", rc)
        return None, d_pooled, dW, db


class _LMPredictionHead(nn.Module):   # LxmertLMPredictionHead (HF:597-607)
    def __init__(self, dims: LxmertDims, embedding_weights: nn.Parameter):
        super().__init__()
        self.transform = _Transform(dims.hidden)
        self.decoder = nn.Linear(dims.hidden, dims.vocab, bias=False)
        self.decoder.weight = embedding_weights
        self.bias = nn.Parameter(torch.zeros(dims.vocab))


class B200LxmertPreTrainingHeads(nn.Module):
    def __init__(self, dims: LxmertDims, embedding_weights: nn.Parameter, passes: int = 3,
                 source: Optional[nn.Module] = None):
        super().__init__()
        if source is not None:
            self.predictions, self.seq_relationship = source.predictions, source.seq_relationship
        else:
            self.predictions = _LMPredictionHead(dims, embedding_weights)
            self.seq_relationship = nn.Linear(dims.hidden, 2)
        self._cdims = _lib.XlxDims.from_dims(dims)
        self.compact_rows = True
        self._fused = _FusedHead("lmhead", dims, self.predictions.decoder.weight.shape[0], passes)

    def _params(self):
        p = self.predictions
        return [p.transform.dense.weight, p.transform.dense.bias, p.transform.LayerNorm.weight,
                p.transform.LayerNorm.bias, p.decoder.weight, p.bias]

    def forward(self, sequence_output, pooled_output):
        """HF signature (HF:662-665) → ``(prediction_scores, seq_relationship_score)``; differentiable when gradients
        are enabled (training loops should prefer :meth:`lm_loss` / :meth:`matched_loss`)."""
        params = self._params()
        if _differentiable(sequence_output, params):
            _, scores = _HeadOutputsFn.apply(self._fused, False, True, sequence_output, *params)
        else:
            scores = self._fused.forward(params, sequence_output, want_logits=True)["logits"]
        if _differentiable(pooled_output, [self.seq_relationship.weight, self.seq_relationship.bias]):
            rel = _MatchScoresFn.apply(self, pooled_output, self.seq_relationship.weight, self.seq_relationship.bias)
        else:
            rel = _match_scores(self, pooled_output)
        return scores, rel

    def lm_loss(self, sequence_output, word_labels):
        """``CrossEntropyLoss()(scores.view(-1, vocab), word_labels.view(-1))`` (modeling.py:219-226)."""
        if self.compact_rows:
            sequence_output, word_labels = labelled_rows(sequence_output, word_labels)
        return _HeadLossFn.apply(self._fused, word_labels, None, False, sequence_output, *self._params())[0]

    def matched_loss(self, pooled_output, matched_labels):
        """``CrossEntropyLoss()(seq_relationship(pooled).view(-1, 2), matched_labels)`` (modeling.py:228-235)."""
        return _MatchLossFn.apply(self, matched_labels, pooled_output, self.seq_relationship.weight,
                                  self.seq_relationship.bias)


class B200LxmertVisualAnswerHead(nn.Module):
    """HF ``LxmertVisualAnswerHead`` (HF:610-623) under its own parameter names ``logit_fc.{0,2,3}.*``:
    ``Linear(H, 2H) → GeLU → LayerNorm(2H, eps 1e-12) → Linear(2H, num_labels)`` on the pooled output.

    ``forward(hidden_states)`` → answer scores (differentiable); ``loss(pooled, labels)`` → ``(CrossEntropyLoss(),
    answer_score.max(1) ids)`` fused (modeling.py:286-299)."""

    def __init__(self, dims: LxmertDims, num_labels: int, passes: int = 3, source: Optional[nn.Module] = None):
        super().__init__()
        if source is not None:
            self.logit_fc = source.logit_fc
            num_labels = source.logit_fc[3].out_features
        else:
            H = dims.hidden
            self.logit_fc = nn.Sequential(nn.Linear(H, 2 * H), nn.GELU(), nn.LayerNorm(2 * H, eps=1e-12),
                                          nn.Linear(2 * H, num_labels))
        self.num_labels = num_labels
        self._fused = _FusedHead("qahead", dims, num_labels, passes)

    def _params(self):
        f = self.logit_fc
        return [f[0].weight, f[0].bias, f[2].weight, f[2].bias, f[3].weight, f[3].bias]

    def forward(self, hidden_states):
        params = self._params()
        if _differentiable(hidden_states, params):
            return _HeadOutputsFn.apply(self._fused, False, True, hidden_states, *params)[1]
        return self._fused.forward(params, hidden_states, want_logits=True)["logits"]

    def loss(self, pooled_output, qa_labels):
        loss, _, pred = _HeadLossFn.apply(self._fused, qa_labels, None, True, pooled_output, *self._params())
        return loss, pred
