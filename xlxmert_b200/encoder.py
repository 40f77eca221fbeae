"""Drop-in replacement for HF ``LxmertEncoder`` (``transformers/models/lxmert/modeling_lxmert.py:487-565``),
the module the reference reaches as ``self.bert.encoder`` (``x-lxmert/src/lxrt/modeling.py:80,195-206``).

``B200LxmertEncoder`` owns the *same* ``nn.Parameter`` objects under the same names as the module it
replaces (``visn_fc.*``, ``layer.N.*``, ``r_layers.N.*``, ``x_layers.N.*``), so state dicts, optimisers and
DDP see no difference; only ``forward`` changes: it hands raw device pointers to
``xlx_encoder_fwd`` / ``xlx_encoder_bwd`` (``include/xlxmert_b200.h``) through a ``torch.autograd.Function``.
There is no PyTorch fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .config import LxmertDims
from .params import encoder_param_names


def dims_from_hf_config(cfg) -> LxmertDims:
    return LxmertDims(hidden=cfg.hidden_size, heads=cfg.num_attention_heads,
                      intermediate=cfg.intermediate_size, feat_dim=cfg.visual_feat_dim,
                      pos_dim=cfg.visual_pos_dim, l_layers=cfg.l_layers, r_layers=cfg.r_layers,
                      x_layers=cfg.x_layers, vocab=cfg.vocab_size, max_pos=cfg.max_position_embeddings,
                      type_vocab=cfg.type_vocab_size, ln_eps=1e-12,
                      hidden_dropout=float(getattr(cfg, "hidden_dropout_prob", 0.0)),
                      attention_dropout=float(getattr(cfg, "attention_probs_dropout_prob", 0.0)))


# ---- a parameter skeleton with the HF names (used when no HF module is at hand) --------------------

class _Att(nn.Module):            # LxmertAttention (HF:217-236)
    def __init__(self, H):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class _AttOut(nn.Module):         # LxmertAttentionOutput / LxmertOutput (HF:277-288, 339-350)
    def __init__(self, K, H):
        super().__init__()
        self.dense = nn.Linear(K, H)
        self.LayerNorm = nn.LayerNorm(H, eps=1e-12)


class _SelfAttLayer(nn.Module):   # LxmertSelfAttentionLayer (HF:306-324)
    def __init__(self, H):
        super().__init__()
        self.self = _Att(H)
        self.output = _AttOut(H, H)


class _CrossAttLayer(nn.Module):  # LxmertCrossAttentionLayer (HF:291-303)
    def __init__(self, H):
        super().__init__()
        self.att = _Att(H)
        self.output = _AttOut(H, H)


class _Inter(nn.Module):          # LxmertIntermediate (HF:327-336)
    def __init__(self, H, I):
        super().__init__()
        self.dense = nn.Linear(H, I)


class _Layer(nn.Module):          # LxmertLayer (HF:353-366)
    def __init__(self, H, I):
        super().__init__()
        self.attention = _SelfAttLayer(H)
        self.intermediate = _Inter(H, I)
        self.output = _AttOut(I, H)


class _XLayer(nn.Module):         # LxmertXLayer (HF:369-384)
    def __init__(self, H, I):
        super().__init__()
        self.visual_attention = _CrossAttLayer(H)
        self.lang_self_att = _SelfAttLayer(H)
        self.visn_self_att = _SelfAttLayer(H)
        self.lang_inter, self.lang_output = _Inter(H, I), _AttOut(I, H)
        self.visn_inter, self.visn_output = _Inter(H, I), _AttOut(I, H)


class _VisnFc(nn.Module):         # LxmertVisualFeatureEncoder (HF:460-474)
    def __init__(self, H, F, P):
        super().__init__()
        self.visn_fc = nn.Linear(F, H)
        self.visn_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.box_fc = nn.Linear(P, H)
        self.box_layer_norm = nn.LayerNorm(H, eps=1e-12)


class _EncoderSkeleton(nn.Module):
    def __init__(self, d: LxmertDims):
        super().__init__()
        self.visn_fc = _VisnFc(d.hidden, d.feat_dim, d.pos_dim)
        self.layer = nn.ModuleList([_Layer(d.hidden, d.intermediate) for _ in range(d.l_layers)])
        self.x_layers = nn.ModuleList([_XLayer(d.hidden, d.intermediate) for _ in range(d.x_layers)])
        self.r_layers = nn.ModuleList([_Layer(d.hidden, d.intermediate) for _ in range(d.r_layers)])


# ---- autograd bridge -------------------------------------------------------------------------------
# Backward stage masks (include/xlxmert_b200.h: XLX_BWD_CROSS = 1, _VISION = 2, _LANGUAGE = 4, _VISN_FC = 8) in the order the
# data-parallel backward issues them: the cross-modality layers first, then everything below them in ONE call so that the
# language stack runs beside the vision stack on the library's second stream.
#: encoders whose stage-wise all-reduces are still in flight (defer_sync_wait); see finish_pending_gradient_sync
_PENDING_SYNC: list = []


def finish_pending_gradient_sync() -> None:
    """Make the current stream wait for every all-reduce an encoder backward left in flight (and apply the 1/world scale
    of backends without an AVG reduction).  Idempotent; ``parallel.allreduce_gradients`` and ``B200AdamW.step`` call it."""
    while _PENDING_SYNC:
        enc = _PENDING_SYNC.pop()
        pending = enc._pending_sync
        if pending is None:
            continue
        works, scale, arena = pending
        for w in works:
            w.wait()
        if scale is not None:
            arena.mul_(scale)
        enc._pending_sync = None


_BWD_STAGES = (1, 2 | 4 | 8)
_BWD_ALL = 15
#: XLX_EARLY_LANGUAGE_REDUCE=1: start the language range's all-reduce while the vision stack of the second backward
#: call still computes.  Off by default — measured at N = 2 (profiles/r02_ab_early_language_reduce.log) it does not pay:
#: the reduce kernels take SMs from the overlapped GEMMs and the tail after the backward is not volume-bound.
_NO_EARLY_LANGUAGE_REDUCE = not bool(int(__import__("os").environ.get("XLX_EARLY_LANGUAGE_REDUCE", "0")))


def _dist_active(group) -> bool:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return False
    return dist.get_world_size(None if group is True else group) > 1


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _EncoderFn(torch.autograd.Function):
    """(lang_in, visual_feats, *params) → (lang_out, vis_out, lang_hidden, vis_hidden)."""

    @staticmethod
    def forward(ctx, enc: "B200LxmertEncoder", lang_mask, visual_pos, vis_mask, want_hidden: bool, training: bool,
                drop, want_attn: bool, lang_in, visual_feats, *params):
        lib = _lib.load()
        d = enc.dims
        B, L, H = lang_in.shape
        V = visual_feats.shape[1]
        dev = lang_in.device
        # `training` comes from the caller: inside Function.forward grad mode is always off and needs_input_grad is
        # True for every parameter even under torch.no_grad(), so neither can tell inference from training
        lang_in = lang_in.contiguous().float()
        visual_feats = visual_feats.contiguous().float()
        visual_pos = visual_pos.contiguous().float()
        prep, parr = enc._prepared(params, force=training)
        ws_bytes = lib.xlx_encoder_workspace_bytes(C.byref(enc._cdims), B, L, V, int(training))
        if ws_bytes == 0:
            raise _lib.XlxError("xlx_encoder_workspace_bytes", -22 if max(L, V) > 64 else -20)
        ws = enc._workspace(ws_bytes, dev, training)
        lang_out = torch.empty(B, L, H, device=dev, dtype=torch.float32)
        vis_out = torch.empty(B, V, H, device=dev, dtype=torch.float32)
        n_l, n_v = d.l_layers + d.x_layers, d.r_layers + d.x_layers
        lang_hidden = torch.empty(n_l, B, L, H, device=dev, dtype=torch.float32) if want_hidden else None
        vis_hidden = torch.empty(n_v, B, V, H, device=dev, dtype=torch.float32) if want_hidden else None
        rc = lib.xlx_encoder_fwd(C.byref(enc._cdims), parr, prep.data_ptr(), B, L, V, lang_in.data_ptr(),
                                 _ptr(lang_mask), visual_feats.data_ptr(), visual_pos.data_ptr(), _ptr(vis_mask),
                                 lang_out.data_ptr(), vis_out.data_ptr(), _ptr(lang_hidden), _ptr(vis_hidden),
                                 ws.data_ptr(), ws_bytes, int(training), enc.passes,
                                 None if drop is None else C.byref(drop), _stream_ptr())
        _lib.check("xlx_encoder_fwd", rc)
        ctx.drop = drop
        # an output the loss does not read must arrive in backward as None (not as a tensor of zeros): it decides which
        # blocks of the last cross-modality layer are part of the graph at all
        ctx.set_materialize_grads(False)
        if training:
            ctx.enc, ctx.ws, ctx.ws_bytes, ctx.shape = enc, ws, ws_bytes, (B, L, V)
            ctx.prep, ctx.parr, ctx.params = prep, parr, params
            ctx.visual_pos = visual_pos
            ctx.need_dfeats = visual_feats.requires_grad
        outs = [lang_out, vis_out]
        if want_hidden:
            ctx.mark_non_differentiable(lang_hidden, vis_hidden)
            outs += [lang_hidden, vis_hidden]
        else:
            outs += [None, None]
        if want_attn:
            # HF:506-565 — language_attentions (layer.*), vision_attentions (r_layers.*), cross_encoder_attentions (the
            # language-query half of every x_layer's cross attention): copies of the probabilities the training plan
            # saves for its backward
            def probs(blk, Sq, Sk):
                off = lib.xlx_encoder_probs_offset(C.byref(enc._cdims), B, L, V, blk, 0)
                if off < 0:
                    raise _lib.XlxError("xlx_encoder_probs_offset", -1)
                n = B * d.heads * Sq * Sk
                return ws[off:off + 4 * n].view(torch.float32).view(B, d.heads, Sq, Sk).clone()
            la = [probs(i, L, L) for i in range(d.l_layers)]
            va = [probs(d.l_layers + i, V, V) for i in range(d.r_layers)]
            xa = [probs(d.l_layers + d.r_layers + 3 * k, L, V) for k in range(d.x_layers)]
            ctx.mark_non_differentiable(*la, *va, *xa)
            ctx.n_attn = len(la) + len(va) + len(xa)
            outs += la + va + xa
        else:
            ctx.n_attn = 0
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_lang_out, d_vis_out, _dlh, _dvh, *_dattn):
        lib = _lib.load()
        enc = ctx.enc
        B, L, V = ctx.shape
        d = enc.dims
        dev = ctx.ws.device
        if d_lang_out is not None:
            d_lang_out = d_lang_out.contiguous().float()
        if d_vis_out is not None:
            d_vis_out = d_vis_out.contiguous().float()
        d_lang_in = torch.empty(B, L, d.hidden, device=dev, dtype=torch.float32)
        d_feats = torch.empty(B, V, d.feat_dim, device=dev, dtype=torch.float32) if ctx.need_dfeats else None
        grads = enc._grad_arena(dev)
        enc.last_grad_arena = grads
        enc.arena_reduced = False

        def run(stages, lang_done=None):
            rc = lib.xlx_encoder_bwd(C.byref(enc._cdims), ctx.parr, ctx.prep.data_ptr(), B, L, V,
                                     ctx.visual_pos.data_ptr(), _ptr(d_lang_out), _ptr(d_vis_out),
                                     d_lang_in.data_ptr(), _ptr(d_feats), grads.data_ptr(), ctx.ws.data_ptr(),
                                     ctx.ws_bytes, enc.passes, stages,
                                     None if ctx.drop is None else C.byref(ctx.drop),
                                     None if lang_done is None else lang_done.cuda_event, _stream_ptr())
            _lib.check("xlx_encoder_bwd", rc)

        group = enc.grad_sync_group
        if group is None or not _dist_active(group):
            run(_BWD_ALL)
        else:
            # data parallel: all-reduce each stage's slice of the arena while the next stage computes
            import torch.distributed as dist
            world = dist.get_world_size(None if group is True else group)
            pg = None if group is True else group
            avg = dist.get_backend(pg) == "nccl"
            works = []
            op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM

            def reduce_range(off, n):
                if n:
                    works.append(dist.all_reduce(grads[off:off + n], op=op, group=pg, async_op=True))
            early = grads.is_cuda and _BWD_STAGES == (1, 14) and not _NO_EARLY_LANGUAGE_REDUCE
            for stage in _BWD_STAGES:
                if early and (stage & 4) and (stage & ~4):
                    # The call runs the language stack beside the vision stack and finishes it first.  The library
                    # records `lang_done` the moment the language range of the arena is complete; an auxiliary stream
                    # that waits on nothing else issues that range's all-reduce, so it overlaps the rest of the call
                    # instead of queueing behind it.
                    lang_done = enc.__dict__.setdefault("_lang_done_event", torch.cuda.Event())
                    aux = enc.__dict__.setdefault("_aux_stream", torch.cuda.Stream(dev))
                    cur = torch.cuda.current_stream(dev)
                    lang_done.record(cur)            # creates the CUDA event; the library re-records it
                    run(stage, lang_done)
                    l_off, l_n = enc._stage_range(4)
                    aux.wait_event(lang_done)
                    with torch.cuda.stream(aux):
                        reduce_range(l_off, l_n)
                    grads.record_stream(aux)
                    for bit in (2, 8):               # vision stack, visual feature encoder: after the call, as before
                        reduce_range(*enc._stage_range(stage & bit)) if stage & bit else None
                else:
                    run(stage)
                    reduce_range(*enc._stage_range(stage))
            enc.arena_reduced = True
            if enc.defer_sync_wait:
                # the caller's stream does not wait here: the rest of the backward (embeddings, heads' leaf gradients) and
                # the packing of the parameters outside the encoder overlap the last range's reduce;
                # parallel.finish_overlapped_sync (called by allreduce_gradients and B200AdamW.step) waits
                enc._pending_sync = (works, None if avg else 1.0 / world, grads)
                _PENDING_SYNC.append(enc)
            else:
                for w in works:
                    w.wait()
                if not avg:
                    grads.mul_(1.0 / world)
        enc._release_workspace(ctx.ws)
        # the last cross-modality layer's self-attention + FFN of a modality whose output got no upstream gradient are
        # outside the graph: no gradient (None), exactly what the reference's autograd reports
        unused = enc._unused_slots(d_lang_out is None, d_vis_out is None)
        pgrads = []
        for i, (p, (off, n)) in enumerate(zip(ctx.params, enc._grad_slices)):
            pgrads.append(grads[off:off + n].view(p.shape) if (p.requires_grad and i not in unused) else None)
        ctx.ws = None
        return (None, None, None, None, None, None, None, None, d_lang_in, d_feats, *pgrads)


class B200LxmertEncoder(nn.Module):
    """``LxmertEncoder`` with the same parameters and ``forward`` signature, running on sm_100a kernels.

    ``passes``: 3 = bf16x3 split GEMMs (fp32-class accuracy, default); 1 = single bf16 pass.
    ``output_hidden_states``: when False (default) the returned hidden-state tuples hold only the final
    state (the reference never asks for the others); set True to get every layer's output like HF.
    """

    def __init__(self, source: Optional[nn.Module] = None, dims: Optional[LxmertDims] = None, passes: int = 3,
                 output_hidden_states: bool = False):
        super().__init__()
        if source is None:
            if dims is None:
                raise ValueError("need either a source LxmertEncoder or dims")
            source = _EncoderSkeleton(dims)
        elif dims is None:
            dims = dims_from_hf_config(source.config)
        self.visn_fc = source.visn_fc
        self.layer = source.layer
        self.x_layers = source.x_layers
        self.r_layers = source.r_layers
        self.config = getattr(source, "config", None)
        self.num_l_layers, self.num_x_layers, self.num_r_layers = dims.l_layers, dims.x_layers, dims.r_layers
        self.dims = dims
        self.defer_sync_wait = False      # see parallel.enable_overlapped_gradient_sync(defer_wait=True)
        self._pending_sync = None
        self.passes = passes
        self.output_hidden_states = output_hidden_states
        self._names = encoder_param_names(dims)
        self._cdims = _lib.XlxDims.from_dims(dims)
        self._prep = None
        self._prep_key = None
        self._parr = None
        self._ws_pool: List[torch.Tensor] = []
        self._grad_slices = None
        #: flat fp32 arena holding every parameter gradient of the most recent backward (the ``.grad`` tensors
        #: are views into it) — one contiguous buffer for the data-parallel all-reduce (lxmert_pretrain.py:104-106)
        self.last_grad_arena: Optional[torch.Tensor] = None
        #: set to ``True`` (default process group) or a ``ProcessGroup`` to have the backward all-reduce (mean) its
        #: gradient arena stage by stage, overlapped with the remaining backward; ``arena_reduced`` then tells
        #: ``parallel.allreduce_gradients`` that nothing is left to do for this module
        self.grad_sync_group = None
        self.arena_reduced = False
        self._stage_ranges = {}

    # -- parameters in C-ABI slot order
    def _param_list(self) -> List[torch.Tensor]:
        """Parameters in C-ABI slot order.  Walking ``named_parameters()`` costs ≈ 1.5 ms per call, so the list is cached
        together with the ``(module, name)`` each entry came from and re-validated by identity (≈ 30 µs): assigning a
        new ``nn.Parameter`` inside any layer, ``.to()`` / ``.cuda()`` conversions and ``load_state_dict`` are all
        picked up; swapping a whole sub-module object for another one requires ``invalidate_parameter_cache()``."""
        cached = self.__dict__.get("_plist")
        if cached is not None:
            owners, plist = cached
            if all(m._parameters.get(n) is p for (m, n), p in zip(owners, plist)):
                return plist
        mods = dict(self.named_modules())
        owners, plist = [], []
        for full in self._names:
            path, _, leaf = full.rpartition(".")
            m = mods[path]
            owners.append((m, leaf))
            plist.append(m._parameters[leaf])
        self.__dict__["_plist"] = (owners, plist)
        return plist

    def invalidate_parameter_cache(self) -> None:
        self.__dict__.pop("_plist", None)

    def _apply(self, fn, recurse=True):
        self.invalidate_parameter_cache()
        return super()._apply(fn, recurse)

    def _prepared(self, params, force: bool = False):
        """Split-bf16 copies of the weights.  Training (``force``): rebuilt on EVERY forward — the weights change with
        every optimiser step, and the reference's own optimiser (HF ``AdamW`` 4.1.1, lxmert_pretrain.py:114) updates
        through ``p.data.add_()``, which leaves no trace in the parameters' version counters, so no cache key could
        notice it (one batched launch, ≈ 0.4 ms of a 45 ms step).  Inference: refreshed when a parameter's
        ``(data_ptr, _version)`` changed (``load_state_dict``, in-place ops under ``no_grad``); after an update through
        ``.data`` call :meth:`invalidate_prepared`."""
        lib = _lib.load()
        key = None if force else tuple((p.data_ptr(), p._version) for p in params)
        dev = params[0].device
        if self._prep is None or self._prep.device != dev:
            nbytes = lib.xlx_encoder_prep_bytes(C.byref(self._cdims))
            self._prep = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._prep_key = None
        if force or key != self._prep_key:
            for p in params:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("B200LxmertEncoder parameters must be contiguous fp32")
            self._parr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            rc = lib.xlx_encoder_prepare(C.byref(self._cdims), self._parr, self._prep.data_ptr(), _stream_ptr())
            _lib.check("xlx_encoder_prepare", rc)
            self._prep_key = key
        return self._prep, self._parr

    def invalidate_prepared(self) -> None:
        """Force the split-bf16 weight copies to be rebuilt on the next forward (after an in-place update
        that bypassed the tensors' version counters, e.g. a fused optimiser writing through raw pointers)."""
        self._prep_key = None

    def _workspace(self, nbytes: int, dev, training: bool) -> torch.Tensor:
        for i, t in enumerate(self._ws_pool):
            if t.numel() >= nbytes and t.device == dev:
                self._ws_pool.pop(i)
                break
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if not training:
            self._ws_pool.append(t)        # stream-ordered reuse is safe for inference
        # training: the autograd node owns the workspace (saved activations) and hands it back in backward; a forward
        # whose graph is dropped without a backward simply frees it
        return t

    def _release_workspace(self, t: torch.Tensor) -> None:
        if len(self._ws_pool) < 2:
            self._ws_pool.append(t)

    def _stage_range(self, stages: int):
        """Element range of the gradient arena completed by a mask of backward stages (their ranges must be adjacent)."""
        if stages not in self._stage_ranges:
            spans = []
            for bit in (1, 2, 4, 8):
                if stages & bit:
                    off, n = C.c_int64(), C.c_int64()
                    _lib.check("xlx_encoder_grad_stage_range",
                               _lib.load().xlx_encoder_grad_stage_range(C.byref(self._cdims), bit, C.byref(off),
                                                                        C.byref(n)))
                    if n.value:
                        spans.append((off.value, n.value))
            spans.sort()
            for (o0, n0), (o1, _) in zip(spans, spans[1:]):
                if o0 + n0 != o1:
                    raise ValueError(f"backward stages {stages:#x} do not cover one contiguous arena range")
            self._stage_ranges[stages] = (spans[0][0], sum(n for _, n in spans)) if spans else (0, 0)
        return self._stage_ranges[stages]

    def _unused_slots(self, no_lang_grad: bool, no_vis_grad: bool):
        """C-ABI slots that ``xlx_encoder_bwd`` leaves unwritten when an output has no upstream gradient (see the
        header): slot layout of an x-layer = ATT(visual_attention) 10, ATT(lang_self_att) 10, ATT(visn_self_att) 10,
        FFN(lang) 6, FFN(visn) 6."""
        d = self.dims
        if d.x_layers == 0 or not (no_lang_grad or no_vis_grad):
            return frozenset()
        base = 8 + (d.l_layers + d.r_layers) * 16 + (d.x_layers - 1) * 42
        out = set()
        if no_lang_grad:
            out.update(range(base + 10, base + 20))
            out.update(range(base + 30, base + 36))
        if no_vis_grad:
            out.update(range(base + 20, base + 30))
            out.update(range(base + 36, base + 42))
        return frozenset(out)

    def _grad_arena(self, dev) -> torch.Tensor:
        lib = _lib.load()
        if self._grad_slices is None:
            n = lib.xlx_encoder_num_params(C.byref(self._cdims))
            self._grad_slices = [(lib.xlx_encoder_grad_offset(C.byref(self._cdims), i),
                                  lib.xlx_encoder_param_elems(C.byref(self._cdims), i)) for i in range(n)]
        total = lib.xlx_encoder_grad_elems(C.byref(self._cdims))
        # The returned gradients are views into the arena, so an arena can only be recycled once nobody holds such a
        # view any more (p.grad = None / zero_grad(set_to_none=True) of the previous step).  Recycling matters: a fresh
        # 0.8 GB request every step fragments PyTorch's caching allocator until it falls back to cudaMalloc, which
        # stalls the host for tens of milliseconds in the middle of a backward.
        pool = self.__dict__.setdefault("_arena_pool", [])
        use_count = getattr(torch._C, "_storage_Use_Count", None)
        if use_count is not None:
            for t in pool:
                if t.device == dev and t.numel() == total:
                    st = t.untyped_storage()
                    if use_count(st._cdata) <= 2:      # the pool's tensor + the wrapper just made
                        return t
        t = torch.empty(total, dtype=torch.float32, device=dev)
        if use_count is not None:
            pool[:] = [a for a in pool if a.device == dev and a.numel() == total][-2:] + [t]
        return t

    # -- partial passes (inference): the language layers never see the visual stream (HF:524-529), so a caller that
    #    re-runs the encoder with the same text (the sampler, tasks/imggen_model.py:199-243) can compute them once
    def _subpass(self, which: str) -> "B200LxmertEncoder":
        subs = self.__dict__.setdefault("_subs", {})         # plain dict: not registered as sub-modules
        if which not in subs:
            from dataclasses import replace
            d = replace(self.dims, r_layers=0, x_layers=0) if which == "lang" else replace(self.dims, l_layers=0)
            subs[which] = B200LxmertEncoder(self, dims=d, passes=self.passes)
        sub = subs[which]
        sub.passes = self.passes
        return sub

    @torch.no_grad()
    def language_stack(self, lang_feats, lang_attention_mask):
        """Output of the ``layer`` (language-only) stack, ``[B, L, H]`` — feed it back through ``forward(...,
        language_stack=...)``.  Inference only."""
        B = lang_feats.shape[0]
        dev = lang_feats.device
        feats = torch.zeros(B, 1, self.dims.feat_dim, device=dev)
        pos = torch.zeros(B, 1, self.dims.pos_dim, device=dev)
        (_, _), (ls, _), _ = self._subpass("lang")(lang_feats, lang_attention_mask, feats, pos)
        return ls[-1]

    def forward(self, lang_feats, lang_attention_mask, visual_feats, visual_pos, visual_attention_mask=None,
                output_attentions=None, language_stack=None):
        if language_stack is not None:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                raise RuntimeError("language_stack= is an inference-only shortcut; wrap the call in torch.no_grad()")
            return self._subpass("rest")(language_stack, lang_attention_mask, visual_feats, visual_pos,
                                         visual_attention_mask)
        if not lang_feats.is_cuda:
            raise RuntimeError("B200LxmertEncoder runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, L, _ = lang_feats.shape
        V = visual_feats.shape[1]

        def flat_mask(m, S):
            if m is None:
                return None
            return m.to(torch.float32).expand(B, 1, 1, S).reshape(B, S).contiguous()

        lmask = flat_mask(lang_attention_mask, L)
        vmask = flat_mask(visual_attention_mask, V)
        params = self._param_list()
        training = torch.is_grad_enabled() and (lang_feats.requires_grad or visual_feats.requires_grad
                                                or any(p.requires_grad for p in params))
        # nn.Dropout follows module.training, not grad mode (HF:474,282,344,236): a train()-mode forward drops even
        # under no_grad, and takes the training plan for it
        drop = _lib.step_dropout(self, self.dims)
        want_attn = bool(output_attentions)
        # the probabilities only exist in the training plan (saved for the backward); note that HF returns them AFTER
        # dropout in train() mode (HF:262-274) — here they are always the undropped soft-max
        training = training or drop is not None or want_attn
        res = _EncoderFn.apply(self, lmask, visual_pos, vmask, self.output_hidden_states, training, drop, want_attn,
                               lang_feats, visual_feats, *params)
        lang_out, vis_out, lh, vh = res[:4]
        if self.output_hidden_states:
            n_l, n_v = lh.shape[0], vh.shape[0]
            lang_states = tuple(lh[i] for i in range(n_l - 1)) + (lang_out,)
            vis_states = tuple(vh[i] for i in range(n_v - 1)) + (vis_out,)
        else:
            lang_states, vis_states = (lang_out,), (vis_out,)
        if not want_attn:
            return ((vis_states, None), (lang_states, None), None)
        d = self.dims
        att = res[4:]
        la, va, xa = att[:d.l_layers], att[d.l_layers:d.l_layers + d.r_layers], att[d.l_layers + d.r_layers:]
        return ((vis_states, tuple(va)), (lang_states, tuple(la)), tuple(xa))


def accelerate(model: nn.Module, passes: int = 3, output_hidden_states: bool = False) -> nn.Module:
    """Swap every HF ``LxmertEncoder`` inside ``model`` for a ``B200LxmertEncoder`` sharing its parameters
    (SURVEY.md §8b "Recommended interposition").  Works on ``XLxmertForPretraining`` (``.bert.encoder``),
    bare ``LxmertModel`` (``.encoder``) and the fine-tune wrappers alike because the swap is per instance."""
    for name, child in list(model.named_children()):
        if type(child).__name__ == "LxmertEncoder":
            setattr(model, name, B200LxmertEncoder(child, passes=passes, output_hidden_states=output_hidden_states))
        else:
            accelerate(child, passes=passes, output_hidden_states=output_hidden_states)
    return model
