"""Pre-training heads on the sm_100a library, under the reference's module and parameter names.

* ``B200LxmertVisualObjHead`` — ``lxrt.modeling.LxmertVisualObjHead`` (``x-lxmert/src/lxrt/modeling.py:8-53``,
  cluster mode): ``forward(hidden_states, out_keys=[])`` → ``{'feat': …, 'obj': …}``; plus the two fused entry
  points the callers actually need: ``loss(hidden, obj_labels)`` (head + ``CrossEntropyLoss``,
  ``modeling.py:244-258``) and ``predict(hidden)`` (head + ``softmax(2).max(2)``, ``tasks/imggen_model.py:228-235``).
* ``B200LxmertPreTrainingHeads`` — HF ``LxmertPreTrainingHeads`` (HF ``modeling_lxmert.py:656-665``) with the
  4.1.1 constructor the reference uses (``modeling.py:86``: decoder weight tied to the word embeddings).

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

    def forward(self, params, hidden, labels=None, want_feat=False, want_logits=False, want_pred=False):
        """→ dict(feat, logits, loss, pred_prob, pred_id, ws, nws)"""
        if not hidden.is_cuda:
            raise RuntimeError("the prediction heads run on CUDA (sm_100a) only; there is no CPU fallback")
        d = self.dims
        lead = hidden.shape[:-1]
        h2 = hidden.reshape(-1, d.hidden).contiguous().float()
        M, dev = h2.shape[0], h2.device
        prep = self.prepared(params, force=labels is not None)      # a loss call is a training step: weights just changed
        if labels is None:      # inference (sampler loop): one grow-only workspace, stream-ordered reuse
            nws = self.fn("workspace_bytes")(C.byref(self.cdims), self.classes, M)
            ws = getattr(self, "_ws_infer", None)
            if ws is None or ws.numel() < nws or ws.device != dev:
                ws = self._ws_infer = torch.empty(nws, dtype=torch.uint8, device=dev)
        else:                   # training: the workspace carries the saved activations to the backward
            ws, nws = self.workspace(M, dev)
        out = dict(ws=ws, nws=nws, prep=prep, M=M, feat=None, logits=None, loss=None, pred_prob=None, pred_id=None)
        if labels is not None:
            labels = labels.reshape(-1).contiguous()
            out["loss"] = torch.empty((), device=dev, dtype=torch.float32)
            out["labels"] = labels
        if want_logits:
            out["logits"] = torch.empty(*lead, self.classes, device=dev, dtype=torch.float32)
        if self.kind == "objhead":
            if want_feat:
                out["feat"] = torch.empty(*lead, d.feat_dim, device=dev, dtype=torch.float32)
            if want_pred:
                out["pred_prob"] = torch.empty(*lead, device=dev, dtype=torch.float32)
                out["pred_id"] = torch.empty(*lead, device=dev, dtype=torch.int64)
            rc = self.fn("fwd")(C.byref(self.cdims), self.classes, _parr(params), prep.data_ptr(), M, h2.data_ptr(),
                                _ptr(labels), _ptr(out["feat"]), _ptr(out["logits"]), _ptr(out["loss"]),
                                _ptr(out["pred_prob"]), _ptr(out["pred_id"]), ws.data_ptr(), nws, self.passes,
                                _stream())
        else:
            rc = self.fn("fwd")(C.byref(self.cdims), self.classes, _parr(params), prep.data_ptr(), M, h2.data_ptr(),
                                _ptr(labels), _ptr(out["logits"]), _ptr(out["loss"]), ws.data_ptr(), nws,
                                self.passes, _stream())
        _lib.check(f"xlx_{self.kind}_fwd", rc)
        return out

    def backward(self, params, saved, d_loss, hidden_shape):
        dev = d_loss.device
        d_loss = d_loss.contiguous().float()
        d_hidden = torch.empty(hidden_shape, device=dev, dtype=torch.float32)
        frozen = 6 if self.kind == "objhead" else -1
        grads = [None if i == frozen else torch.empty_like(p) for i, p in enumerate(params)]
        rc = self.fn("bwd")(C.byref(self.cdims), self.classes, _parr(params), saved["prep"].data_ptr(), saved["M"],
                            saved["labels"].data_ptr(), d_loss.data_ptr(), d_hidden.data_ptr(), _parr(grads),
                            saved["ws"].data_ptr(), saved["nws"], self.passes, _stream())
        _lib.check(f"xlx_{self.kind}_bwd", rc)
        return d_hidden, grads


class _HeadLossFn(torch.autograd.Function):
    """(hidden, *params) → CrossEntropyLoss(head(hidden), labels) as a 0-d tensor."""

    @staticmethod
    def forward(ctx, fused: _FusedHead, labels, hidden, *params):
        out = fused.forward(list(params), hidden, labels=labels)
        if any(ctx.needs_input_grad):
            # everything but the returned loss: an output stored on ctx would tie the workspace (GBs of logits) into a
            # reference cycle that only the cyclic garbage collector frees
            saved = {k: v for k, v in out.items() if k != "loss"}
            ctx.fused, ctx.saved, ctx.params, ctx.hshape = fused, saved, params, hidden.shape
        return out["loss"]

    @staticmethod
    def backward(ctx, d_loss):
        d_hidden, grads = ctx.fused.backward(list(ctx.params), ctx.saved, d_loss, ctx.hshape)
        ctx.saved = None
        return (None, None, d_hidden, *[g if (g is not None and p.requires_grad) else None
                                         for g, p in zip(grads, ctx.params)])


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


def labelled_rows(hidden, labels):
    """Drop the rows CrossEntropyLoss ignores (label −100, modeling.py:99,253-256) before a masked-prediction head:
    same loss, same gradients, none of the head's GEMM work for rows that cannot contribute.  Costs one 4-byte
    device→host read (the row count sizes the head's GEMMs).  Returns ``(hidden, labels)`` unchanged when every row
    is labelled or none is (the all-ignored loss is NaN like the reference's)."""
    if not hidden.is_cuda:
        raise RuntimeError("the prediction heads run on CUDA (sm_100a) only; there is no CPU fallback")
    H = hidden.shape[-1]
    flat = labels.reshape(-1).contiguous()
    M = flat.numel()
    rows = torch.empty(M, dtype=torch.int64, device=hidden.device)
    count = torch.empty(1, dtype=torch.int32, device=hidden.device)
    lib = _lib.load()
    _lib.check("xlx_labelled_rows", lib.xlx_labelled_rows(flat.data_ptr(), M, -100, rows.data_ptr(), count.data_ptr(),
                                                          _stream()))
    n = int(count.item())
    if n == 0 or n == M:
        return hidden, labels
    h2 = hidden.reshape(M, H)
    if h2.dtype != torch.float32 or not h2.is_contiguous():
        h2 = h2.contiguous().float()
    picked = torch.empty(n, dtype=torch.int64, device=hidden.device)
    _lib.check("xlx_gather_rows", lib.xlx_gather_rows(None, flat.data_ptr(), rows.data_ptr(), n, 0, None,
                                                      picked.data_ptr(), _stream()))
    return _LabelledRowsFn.apply(h2, rows, n), picked


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
        self.visual_losses = {"obj": {"shape": (-1,), "num": C_}}     # --visualLosses obj (pretrain.bash)
        self._fused = _FusedHead("objhead", dims, C_, passes)

    def _params(self) -> List[torch.Tensor]:
        t = self.transform
        return [t.dense.weight, t.dense.bias, t.LayerNorm.weight, t.LayerNorm.bias, self.linear_feat.weight,
                self.linear_feat.bias, self.out_cluster.weight, self.out_cluster.bias]

    def forward(self, hidden_states, out_keys=[]):
        """Reference signature (modeling.py:38).  Returns plain tensors (no autograd graph): training goes
        through :meth:`loss`, which fuses the cross-entropy and has a native backward."""
        if torch.is_grad_enabled() and hidden_states.requires_grad:
            raise RuntimeError("B200LxmertVisualObjHead.forward returns non-differentiable logits; "
                               "use .loss(hidden, obj_labels) for training")
        keys = set(self.visual_losses) | set(out_keys)
        out = self._fused.forward(self._params(), hidden_states, want_feat="feat" in keys, want_logits="obj" in keys)
        res = {}
        if "feat" in keys:
            res["feat"] = out["feat"]
        if "obj" in keys:
            res["obj"] = out["logits"]
        return res

    def loss(self, hidden_states, obj_labels):
        """``CrossEntropyLoss()(obj_logit.view(B·V, C), obj_label.flatten())`` (modeling.py:253-256), differentiable."""
        if self.compact_rows:
            hidden_states, obj_labels = labelled_rows(hidden_states, obj_labels)
        return _HeadLossFn.apply(self._fused, obj_labels, hidden_states, *self._params())

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
        """HF signature (HF:662-665) → ``(prediction_scores, seq_relationship_score)``, non-differentiable."""
        if torch.is_grad_enabled() and (sequence_output.requires_grad or pooled_output.requires_grad):
            raise RuntimeError("B200LxmertPreTrainingHeads.forward returns non-differentiable scores; "
                               "use .lm_loss / .matched_loss for training")
        out = self._fused.forward(self._params(), sequence_output, want_logits=True)
        lib = _lib.load()
        B = pooled_output.shape[0]
        pooled = pooled_output.contiguous().float()
        rel = torch.empty(B, 2, device=pooled.device, dtype=torch.float32)
        rc = lib.xlx_matchhead_fwd(C.byref(self._cdims), B, pooled.data_ptr(), self.seq_relationship.weight.data_ptr(),
                                   self.seq_relationship.bias.data_ptr(), None, rel.data_ptr(), None, None, _stream())
        _lib.check("xlx_matchhead_fwd", rc)
        return out["logits"], rel

    def lm_loss(self, sequence_output, word_labels):
        """``CrossEntropyLoss()(scores.view(-1, vocab), word_labels.view(-1))`` (modeling.py:219-226)."""
        if self.compact_rows:
            sequence_output, word_labels = labelled_rows(sequence_output, word_labels)
        return _HeadLossFn.apply(self._fused, word_labels, sequence_output, *self._params())

    def matched_loss(self, pooled_output, matched_labels):
        """``CrossEntropyLoss()(seq_relationship(pooled).view(-1, 2), matched_labels)`` (modeling.py:228-235)."""
        return _MatchLossFn.apply(self, matched_labels, pooled_output, self.seq_relationship.weight,
                                  self.seq_relationship.bias)
