"""Input path of the pre-training step: drop-in for the argument building of
``Trainer.forward`` (``x-lxmert/src/pretrain/lxmert_pretrain.py:143-225``).

The reference moves a collate_fn batch (``lxmert_data.py:497-652``) to the GPU with up to 13 separate ``.to(device)``
copies from pageable memory and builds labels / masks with a dozen tiny torch kernels per step.  ``B200PretrainInputs``
packs the step's arrays into ONE pinned host buffer, issues ONE ``cudaMemcpyAsync`` on a copy stream and ONE unpack
kernel (``xlx_pretrain_inputs_unpack``: ``attention_mask = word_id > 0``, the additive mask, ``obj_labels[~vis_mask] =
-100``, …), double-buffered so that step *i + 1* is staged while step *i* computes::

    inputs = B200PretrainInputs(device)
    inputs.stage(batch, task)              # host: pack + enqueue (returns at once)
    out = model(**inputs.kwargs())         # same keyword arguments Trainer.forward passes (lxmert_pretrain.py:202-223)

Clustering mode (``--clustering``, pretrain.bash): ``cluster_id`` instead of ``vis_feats`` as the visual input.
Optional extras of the published defaults: ``qa_labels=True`` (``--taskQA``) packs ``qa_label`` and applies the
``matched`` rule (``lxmert_pretrain.py:184-189``); ``feat_labels=True`` (``--visualLosses obj,feat``) ships
``batch['vis_feats']`` — the regression targets, ``:177-179`` — with one extra copy on the same stream (straight from the
batch tensor when the DataLoader pinned it).  Integer / byte work — bit-exact against the reference's torch statements
(tests/test_inputs.py).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib

TASKS = ("vis_mask", "word_mask", "matched")          # MASK_MODALITY order (lxmert_pretrain.py:794-800)
_WORD_KEY = {"vis_mask": "word_id", "word_mask": "masked_word_id", "matched": "other_word_id"}   # :193-198


def packed_layout(B: int, L: int, V: int):
    """Byte offsets of the seven sections of the packed buffer and its total size (the C ABI owns the layout)."""
    offs, total = (C.c_int64 * 7)(), C.c_int64()
    _lib.check("xlx_pretrain_inputs_layout", _lib.load().xlx_pretrain_inputs_layout(B, L, V, offs, C.byref(total)))
    return list(offs), int(total.value)


class _Slot:
    def __init__(self):
        self.host: Optional[torch.Tensor] = None       # pinned uint8
        self.dev: Optional[torch.Tensor] = None        # packed copy on the device
        self.out: Optional[Dict[str, torch.Tensor]] = None
        self.ready: Optional[torch.cuda.Event] = None  # unpack finished (recorded on the copy stream)
        self.free: Optional[torch.cuda.Event] = None   # consumer enqueued everything that reads `out`
        self.shape = None
        self.task = None
        self.free_recorded = False


class B200PretrainInputs:
    def __init__(self, device=None, depth: int = 2, qa_labels: bool = False, feat_labels: bool = False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("B200PretrainInputs stages batches onto a CUDA device (sm_100a); there is no CPU fallback")
        self._slots: List[_Slot] = [_Slot() for _ in range(max(1, depth))]
        self._next = 0
        self._staged: List[_Slot] = []
        self._copy_stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.qa_labels, self.feat_labels = bool(qa_labels), bool(feat_labels)

    # ---- host side -------------------------------------------------------------------------------------------------
    def _prepare_slot(self, s: _Slot, B: int, L: int, V: int):
        if s.shape == (B, L, V):
            return
        offs, total = packed_layout(B, L, V)
        s.offs, s.total = offs, total
        s.host = torch.empty(total, dtype=torch.uint8).pin_memory()
        dev = self.device
        s.dev = torch.empty(total, dtype=torch.uint8, device=dev)
        s.out = dict(
            word_id=torch.empty(B, L, dtype=torch.int64, device=dev),
            attention_mask=torch.empty(B, L, dtype=torch.bool, device=dev),
            additive_mask=torch.empty(B, L, dtype=torch.float32, device=dev),
            cluster_ids=torch.empty(B, V, dtype=torch.int64, device=dev),
            vis_mask=torch.empty(B, V, dtype=torch.bool, device=dev),
            visual_pos=torch.empty(B, V, 4, dtype=torch.float32, device=dev),
            obj_labels=torch.empty(B, V, dtype=torch.int64, device=dev),
            word_labels=torch.empty(B, L, dtype=torch.int64, device=dev),
            matched_labels=torch.empty(B, dtype=torch.int64, device=dev),
            qa_labels=torch.empty(B, dtype=torch.int64, device=dev))
        s.feat_host = s.feat_dev = None
        s.shape = (B, L, V)
        s.ready = torch.cuda.Event()
        s.free = None

    @staticmethod
    def _put(host: torch.Tensor, off: int, t: torch.Tensor, dtype):
        t = t.detach()
        if t.is_cuda:
            raise TypeError("B200PretrainInputs packs HOST batches (the DataLoader's output)")
        t = t.to(dtype).contiguous()
        n = t.numel() * t.element_size()
        host[off:off + n].copy_(t.view(-1).view(torch.uint8))

    def stage(self, batch: dict, task: str) -> None:
        """Pack ``batch`` (CPU tensors as collate_fn returns them) for ``task`` and enqueue copy + unpack."""
        if task not in TASKS:
            raise ValueError(f"task must be one of {TASKS} (got {task!r})")
        ids = batch[_WORD_KEY[task]]
        B, L = ids.shape
        V = batch["cluster_id"].shape[1]
        s = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if s in self._staged:
            raise RuntimeError("more batches staged than slots: call kwargs() before staging again (or raise depth)")
        self._prepare_slot(s, B, L, V)
        if s.free is not None:             # handed out before: its tensors may still be read by an enqueued step
            if s.free_recorded:
                s.free.synchronize()
            else:                          # the consumer never called done(): fall back to a full device sync
                torch.cuda.synchronize(self.device)
            s.free = None
        o = s.offs
        self._put(s.host, o[0], ids, torch.int64)
        if task == "word_mask":
            self._put(s.host, o[1], batch["word_label"], torch.int64)
        if task == "matched":
            self._put(s.host, o[2], batch["matched_label"], torch.int64)
        self._put(s.host, o[3], batch["cluster_id"], torch.int64)
        self._put(s.host, o[4], batch["vis_mask"], torch.uint8)
        self._put(s.host, o[5], batch["box_position"], torch.float32)
        if self.qa_labels:
            self._put(s.host, o[6], batch["qa_label"], torch.int64)
            if task == "matched" and "matched_label" not in batch:
                raise KeyError("matched_label")
        feats = None
        if self.feat_labels and task == "vis_mask":
            feats = batch["vis_feats"].detach()
            if feats.is_cuda or feats.dtype != torch.float32:
                raise TypeError("vis_feats must be a host float32 tensor [B, V, feat_dim]")
            feats = feats.contiguous()
            if s.feat_dev is None or s.feat_dev.shape != feats.shape:
                s.feat_dev = torch.empty(feats.shape, dtype=torch.float32, device=self.device)
            if not feats.is_pinned():          # pageable DataLoader output: through this slot's pinned staging buffer
                if s.feat_host is None or s.feat_host.shape != feats.shape:
                    s.feat_host = torch.empty(feats.shape, dtype=torch.float32).pin_memory()
                s.feat_host.copy_(feats)
                feats = s.feat_host
        lib = _lib.load()
        out = s.out
        with torch.cuda.stream(self._copy_stream):
            s.dev.copy_(s.host, non_blocking=True)
            rc = lib.xlx_pretrain_inputs_unpack(
                s.dev.data_ptr(), B, L, V, TASKS.index(task), out["word_id"].data_ptr(),
                out["attention_mask"].data_ptr(), out["additive_mask"].data_ptr(), out["cluster_ids"].data_ptr(),
                out["vis_mask"].data_ptr(), out["visual_pos"].data_ptr(), out["obj_labels"].data_ptr(),
                out["word_labels"].data_ptr(), out["matched_labels"].data_ptr(),
                out["qa_labels"].data_ptr() if self.qa_labels else None, self._copy_stream.cuda_stream)
            _lib.check("xlx_pretrain_inputs_unpack", rc)
            if feats is not None:
                s.feat_dev.copy_(feats, non_blocking=True)
            s.ready.record(self._copy_stream)
        s.task = task
        s.has_feats = feats is not None
        self.h2d_bytes = s.total + (feats.numel() * 4 if feats is not None else 0)
        self._staged.append(s)

    # ---- consumer side ---------------------------------------------------------------------------------------------
    def kwargs(self) -> dict:
        """Keyword arguments of ``XLxmertForPretraining.forward`` for the oldest staged batch — the ones
        ``Trainer.forward`` passes (lxmert_pretrain.py:202-223).  The current stream waits for the unpack kernel; the
        host does not."""
        if not self._staged:
            raise RuntimeError("kwargs() without a staged batch")
        s = self._staged.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(s.ready)
        o = s.out
        am = o["attention_mask"]
        am._xlx_additive = o["additive_mask"]          # B200LxmertModel picks the prebuilt additive mask up
        label_dict = {}
        if s.task == "vis_mask":
            label_dict["obj_labels"] = o["obj_labels"]
        elif s.task == "word_mask":
            label_dict["word_labels"] = o["word_labels"]
        else:
            label_dict["matched_labels"] = o["matched_labels"]
        if self.qa_labels:
            label_dict["qa_labels"] = o["qa_labels"]
        if s.has_feats:
            label_dict["feat_labels"] = s.feat_dev
        s.free = torch.cuda.Event()
        s.free_recorded = False
        self._pending_free = s
        return dict(input_ids=o["word_id"], visual_feats=None, visual_pos=o["visual_pos"], attention_mask=am,
                    visual_attention_mask=None, cluster_ids=o["cluster_ids"], vis_mask=o["vis_mask"],
                    token_type_ids=None, return_dict=True, label_dict=label_dict, task=s.task)

    def done(self) -> None:
        """Call after the step that consumed ``kwargs()`` has been enqueued (forward + backward): marks the slot's
        tensors reusable once the device gets there."""
        s = getattr(self, "_pending_free", None)
        if s is not None:
            s.free.record(torch.cuda.current_stream(self.device))
            s.free_recorded = True
            self._pending_free = None
