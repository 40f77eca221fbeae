"""``LxmertEmbeddings`` (HF ``modeling_lxmert.py:179-214``) and ``LxmertPooler`` (HF:568-580) on the sm_100a
library, with the reference's parameter names (``word_embeddings.weight`` …, ``dense.weight`` …).

Dropout (HF:189,212) is not applied: parity runs use p = 0 / ``eval()``; training-mode
dropout is statistical and not reproducible from the fused path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib
from .config import LxmertDims


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _parr(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _EmbeddingsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "B200LxmertEmbeddings", input_ids, token_type_ids, word, pos, typ, ln_w, ln_b):
        lib = _lib.load()
        B, L = input_ids.shape
        H = word.shape[1]
        dev = word.device
        if L > pos.shape[0]:
            raise ValueError(f"sequence length {L} exceeds max_position_embeddings {pos.shape[0]}")
        ids = input_ids.contiguous()
        tt = None if token_type_ids is None else token_type_ids.contiguous()
        params = [word, pos, typ, ln_w, ln_b]
        training = any(ctx.needs_input_grad)
        out = torch.empty(B, L, H, device=dev, dtype=torch.float32)
        save = None
        if training:
            save = torch.empty(lib.xlx_embeddings_save_bytes(C.byref(mod._cdims), B, L), dtype=torch.uint8, device=dev)
        rc = lib.xlx_embeddings_fwd(C.byref(mod._cdims), B, L, ids.data_ptr(), None if tt is None else tt.data_ptr(),
                                    _parr(params), out.data_ptr(), None if save is None else save.data_ptr(),
                                    _stream())
        _lib.check("xlx_embeddings_fwd", rc)
        if training:
            ctx.mod, ctx.ids, ctx.tt, ctx.save, ctx.params = mod, ids, tt, save, params
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        mod, ids, tt, params = ctx.mod, ctx.ids, ctx.tt, ctx.params
        B, L = ids.shape
        dev = d_out.device
        d_out = d_out.contiguous().float()
        grads = [torch.empty_like(p) for p in params]
        nscr = lib.xlx_embeddings_scratch_bytes(C.byref(mod._cdims), B, L)
        scratch = torch.empty(nscr, dtype=torch.uint8, device=dev)
        rc = lib.xlx_embeddings_bwd(C.byref(mod._cdims), B, L, params[0].shape[0], params[1].shape[0],
                                    params[2].shape[0], ids.data_ptr(), None if tt is None else tt.data_ptr(),
                                    _parr(params), ctx.save.data_ptr(), d_out.data_ptr(), _parr(grads),
                                    scratch.data_ptr(), nscr, _stream())
        _lib.check("xlx_embeddings_bwd", rc)
        ctx.save = None
        return (None, None, None, *[g if p.requires_grad else None for g, p in zip(grads, params)])


class B200LxmertEmbeddings(nn.Module):
    def __init__(self, dims: LxmertDims, source: Optional[nn.Module] = None):
        super().__init__()
        H = dims.hidden
        if source is not None:
            self.word_embeddings = source.word_embeddings
            self.position_embeddings = source.position_embeddings
            self.token_type_embeddings = source.token_type_embeddings
            self.LayerNorm = source.LayerNorm
        else:
            self.word_embeddings = nn.Embedding(dims.vocab, H, padding_idx=0)
            self.position_embeddings = nn.Embedding(dims.max_pos, H, padding_idx=0)
            self.token_type_embeddings = nn.Embedding(dims.type_vocab, H, padding_idx=0)
            self.LayerNorm = nn.LayerNorm(H, eps=1e-12)
        self.dims = dims
        self._cdims = _lib.XlxDims.from_dims(dims)

    def forward(self, input_ids, token_type_ids=None, inputs_embeds=None):
        if inputs_embeds is not None:
            raise NotImplementedError("inputs_embeds is not supported by the fused embedding path")
        if not input_ids.is_cuda:
            raise RuntimeError("B200LxmertEmbeddings runs on CUDA (sm_100a) only; there is no CPU fallback")
        return _EmbeddingsFn.apply(self, input_ids, token_type_ids, self.word_embeddings.weight,
                                   self.position_embeddings.weight, self.token_type_embeddings.weight,
                                   self.LayerNorm.weight, self.LayerNorm.bias)


class _PoolerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "B200LxmertPooler", lang_out, W, b):
        lib = _lib.load()
        B, L, H = lang_out.shape
        dev = lang_out.device
        lang_out = lang_out.contiguous().float()
        nws = lib.xlx_pooler_workspace_bytes(C.byref(mod._cdims), B)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        pooled = torch.empty(B, H, device=dev, dtype=torch.float32)
        rc = lib.xlx_pooler_fwd(C.byref(mod._cdims), B, L, lang_out.data_ptr(), W.data_ptr(), b.data_ptr(),
                                pooled.data_ptr(), ws.data_ptr(), nws, mod.passes, _stream())
        _lib.check("xlx_pooler_fwd", rc)
        if any(ctx.needs_input_grad):
            ctx.mod, ctx.ws, ctx.nws, ctx.shape = mod, ws, nws, (B, L, H)
            ctx.save_for_backward(pooled)     # an output kept as a plain ctx attribute would form a reference cycle
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        lib = _lib.load()
        mod = ctx.mod
        B, L, H = ctx.shape
        dev = d_pooled.device
        d_pooled = d_pooled.contiguous().float()
        d_lang = torch.empty(B, L, H, device=dev, dtype=torch.float32)
        dW = torch.empty(H, H, device=dev, dtype=torch.float32)
        db = torch.empty(H, device=dev, dtype=torch.float32)
        (pooled,) = ctx.saved_tensors
        rc = lib.xlx_pooler_bwd(C.byref(mod._cdims), B, L, pooled.data_ptr(), d_pooled.data_ptr(),
                                d_lang.data_ptr(), dW.data_ptr(), db.data_ptr(), ctx.ws.data_ptr(), ctx.nws,
                                mod.passes, _stream())
        _lib.check("xlx_pooler_bwd", rc)
        ctx.ws = None
        return None, d_lang, dW, db


class _Dense(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.dense = nn.Linear(H, H)


class B200LxmertPooler(nn.Module):
    def __init__(self, dims: LxmertDims, passes: int = 3, source: Optional[nn.Module] = None):
        super().__init__()
        self.dense = source.dense if source is not None else nn.Linear(dims.hidden, dims.hidden)
        self.passes = passes
        self._cdims = _lib.XlxDims.from_dims(dims)

    def forward(self, hidden_states):
        if not hidden_states.is_cuda:
            raise RuntimeError("B200LxmertPooler runs on CUDA (sm_100a) only; there is no CPU fallback")
        return _PoolerFn.apply(self, hidden_states, self.dense.weight, self.dense.bias)
