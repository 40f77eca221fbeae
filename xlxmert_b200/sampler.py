"""Text → image sampling model: drop-in for ``ImggenModel`` (``x-lxmert/src/tasks/imggen_model.py:11-257``).

Same attributes (``bert``, ``obj_predict_head``, ``mask_feat``, ``vis_emb``, ``G``), same ``set_visual_embedding`` /
``set_image_generator`` / ``denorm`` and the two samplers with the reference's keywords.  Every encoder pass, the
cluster head with ``softmax(2).max(2)`` fused into the logits GEMM, the loop transitions (confidence ranking, re-masking,
centroid gather — the reference's ``topk`` / ``scatter_`` / ``where`` / ``vis_emb``) and the generator run in
``libxlxmert_b200.so``; the host only computes the schedule, nothing is read back between steps.

``sentences`` may be a list of strings (needs a tokenizer: pass one to the constructor — the reference downloads
``unc-nlp/lxmert-base-uncased``, impossible offline) or an already tokenised ``LongTensor [B, L]``.
"""
from __future__ import annotations

import random
from typing import Optional

import numpy as np
import torch
from torch import nn

from .config import LxmertDims
from .heads import B200LxmertVisualObjHead
from .lxmert import B200LxmertModel
from .synth import box_position


class B200ImggenModel(nn.Module):
    def __init__(self, config, args=None, num_clusters: int = 10000, passes: int = 3, tokenizer=None):
        super().__init__()
        from .encoder import dims_from_hf_config
        dims = config if isinstance(config, LxmertDims) else dims_from_hf_config(config)
        if dims.num_clusters != num_clusters:
            dims = LxmertDims(**{**dims.asdict(), "num_clusters": num_clusters})
        self.dims, self.args = dims, args
        self.config = None if isinstance(config, LxmertDims) else config
        self.bert = B200LxmertModel(dims, passes=passes)
        self.obj_predict_head = B200LxmertVisualObjHead(dims, num_clusters, passes=passes)
        self.mask_feat = nn.Parameter(torch.zeros(dims.feat_dim))
        self.vis_emb: Optional[nn.Embedding] = None
        self.tokenizer = tokenizer
        self.G = None

    def set_visual_embedding(self, centroids):                         # imggen_model.py:29-39
        if isinstance(centroids, np.ndarray):
            centroids = torch.from_numpy(centroids)
        centroids = centroids.to(device=self.mask_feat.device, dtype=torch.float32).contiguous()
        self.vis_emb = nn.Embedding.from_pretrained(centroids, freeze=True)
        self.obj_predict_head.out_cluster.weight = self.vis_emb.weight

    def set_image_generator(self, generator):                          # imggen_model.py:41-42
        self.G = generator

    def denorm(self, x):                                               # imggen_model.py:44-47
        """(-1, 1) => (0, 1)"""
        return ((x + 1) / 2).clamp(0, 1)

    # -- helpers -------------------------------------------------------------------------------------
    def _input_ids(self, sentences, max_text_length):
        dev = self.mask_feat.device
        if torch.is_tensor(sentences):
            return sentences[:, :max_text_length].to(dev)
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer: pass tokenizer= to B200ImggenModel or give token ids [B, L]")
        ids = self.tokenizer(sentences, max_length=max_text_length, truncation=True, return_tensors='pt').input_ids
        return ids.to(dev)

    def _predict(self, input_ids, code, visual_pos, language_stack=None):
        """One pass: LXMERT → cluster head → ``softmax(2).max(2)`` (imggen_model.py:221-235).  The reference re-runs
        the nine language-only layers on the unchanged text at every step; with ``language_stack`` they are reused
        (bit-identical result, tests/test_sampler.py)."""
        out = self.bert(input_ids=input_ids, visual_feats=code, visual_pos=visual_pos, attention_mask=input_ids > 0,
                        language_stack=language_stack)
        return self.obj_predict_head.predict(out[1])

    # -- CUDA-graph replay of the per-step device work -----------------------------------------------------------
    # At sampling batch sizes the step (≈ 150 kernels of a few µs each) is launch-bound; the whole pass — rest of the
    # encoder on the cached language stack → cluster head → softmax-max/arg-max — is captured once per (B, L) into a
    # CUDA graph over static buffers and replayed every step.
    def _graph_predict(self, input_ids, code, visual_pos, language_stack):
        # the weight-prepare kernels run during warm-up, outside the capture: a graph is only valid for the parameter
        # values it was captured with, so the parameters' (data_ptr, version) are part of the key (load_state_dict,
        # set_visual_embedding, fine-tuning → a fresh capture; stale graphs are dropped)
        pkey = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self.vis_emb is not None:
            pkey += ((self.vis_emb.weight.data_ptr(), self.vis_emb.weight._version),)
        graphs = self.__dict__.setdefault("_graphs", {})
        if self.__dict__.get("_graphs_pkey") != pkey:
            graphs.clear()
            self.__dict__["_graphs_pkey"] = pkey
        key = (tuple(input_ids.shape), code.device.index)
        st = graphs.get(key)
        if st is None:
            st = {"code": torch.empty_like(code), "lang": torch.empty_like(language_stack),
                  "ids": input_ids.clone(), "pos": visual_pos.clone()}
            st["code"].copy_(code)
            st["lang"].copy_(language_stack)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside capture: weight prep, lazy function attributes
                for _ in range(2):
                    self._predict(st["ids"], st["code"], st["pos"], st["lang"])
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["prob"], st["id"] = self._predict(st["ids"], st["code"], st["pos"], st["lang"])
            st["graph"] = g
            self._graphs[key] = st
        st["code"].copy_(code)
        st["lang"].copy_(language_stack)
        st["ids"].copy_(input_ids)
        st["graph"].replay()
        return st["prob"].clone(), st["id"].clone()

    def _decode(self, code, B, code_dim, grid_size):
        """Generator → denorm → CPU tensor like the reference (imggen_model.py:157,165).  The device→host copy lands in
        page-locked memory (≈ 3× faster than a pageable ``.cpu()`` for the 25 MB of a 32-image batch) that the caller
        then owns: every call takes a fresh block from torch's caching host allocator — after the first call a free-list
        hit, not a cudaHostAlloc — so nothing is copied a second time on the host."""
        img = self.denorm(self.G(code.permute(0, 2, 1).view(B, code_dim, grid_size, grid_size)))
        out = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        out.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    # -- device-side loop transitions (csrc/sampler.cu) --------------------------------------------------------------
    def _nar_update(self, code, vis_mask, pred_prob, pred_id, n_mask_next):
        """topk(largest=False) + scatter_ + the two torch.where of imggen_model.py:209-218,238-243 in one kernel."""
        from . import _lib
        B, V, F = code.shape
        rc = _lib.load().xlx_sampler_nar_update(
            code.data_ptr(), vis_mask.data_ptr(), None if pred_prob is None else pred_prob.data_ptr(),
            None if pred_id is None else pred_id.data_ptr(), self.vis_emb.weight.data_ptr(), self._mask_feat32().data_ptr(),
            B, V, F, n_mask_next, vis_mask.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check("xlx_sampler_nar_update", rc)

    def _mask_feat32(self):
        mf = self.mask_feat.detach()
        return mf if (mf.dtype == torch.float32 and mf.is_contiguous()) else mf.float().contiguous()

    def _visual_pos(self, B, grid_size, dev):
        """``box_position(grid_size)`` expanded to the batch (imggen_model.py:195), cached on the device: building it
        from numpy on every call costs a pageable host→device copy and a synchronisation."""
        key = (B, grid_size, str(dev))
        cache = self.__dict__.setdefault("_vpos_cache", {})
        if key not in cache:
            cache.clear()
            cache[key] = torch.from_numpy(box_position(grid_size)).unsqueeze(0).expand(B, -1, -1).contiguous().to(dev)
        return cache[key]

    def _check_device(self, input_ids):
        if not input_ids.is_cuda or self.vis_emb is None or not self.vis_emb.weight.is_cuda:
            raise RuntimeError("B200ImggenModel samples on CUDA (sm_100a) only — move the model and its visual "
                               "embedding to the GPU; there is no CPU fallback")

    # -- samplers ------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_image_NAR(self, sentences, max_text_length=20, n_steps=None, return_intermediate=False,
                         return_codes=False, cache_language=True, cuda_graph=False):
        """Mask-predict sampling with linear decay (imggen_model.py:169-257).  The loop body never returns to the host:
        encoder pass → cluster head with the soft-max / arg-max fused into the logits GEMM → one transition kernel
        (re-masking by confidence rank + centroid gather); only the schedule ``n_mask`` is host arithmetic."""
        self.eval()
        input_ids = self._input_ids(sentences, max_text_length)
        self._check_device(input_ids)
        B, dev = input_ids.shape[0], input_ids.device
        grid_size, code_dim = 8, self.dims.feat_dim
        n_grids = grid_size ** 2
        if n_steps is None:
            n_steps = n_grids
        visual_pos = self._visual_pos(B, grid_size, dev)
        intermediate_imgs = []
        pred_prob = pred_code_id = None
        lang = self.bert.language_stack(input_ids, input_ids > 0) if cache_language else None
        code = torch.empty(B, n_grids, code_dim, device=dev, dtype=torch.float32)
        vis_mask = torch.empty(B, n_grids, dtype=torch.uint8, device=dev)
        self._nar_update(code, vis_mask, None, None, n_grids)          # i = 0: every cell masked (:204-206,215-218)
        for i in range(n_steps):
            if cuda_graph and lang is not None:
                pred_prob, pred_code_id = self._graph_predict(input_ids, code, visual_pos, lang)
            else:
                pred_prob, pred_code_id = self._predict(input_ids, code, visual_pos, lang)
            # cells masked in this step take their predicted centroid; the next step's mask = the n_mask least confident
            n_next = int((n_steps - (i + 1)) / n_steps * n_grids) if i + 1 < n_steps else 0
            if return_intermediate:
                # the reference decodes the fully predicted grid of every step, before re-masking
                full = code.clone()
                keep = vis_mask.clone()
                self._nar_update(full, keep, pred_prob, pred_code_id, 0)
                intermediate_imgs.append(self._decode(full, B, code_dim, grid_size))
            self._nar_update(code, vis_mask, pred_prob, pred_code_id, n_next)
        if return_intermediate:
            return intermediate_imgs
        if return_codes:
            return code, pred_prob, pred_code_id
        return self._decode(code, B, code_dim, grid_size)

    @torch.no_grad()
    def sample_image_AR(self, sentences, max_text_length=20, position_random=False, position_TLBR=False,
                        position_confidence=True, n_steps=None, seed=None, return_intermediate=False,
                        return_codes=False, cache_language=True):
        """One grid cell per step (imggen_model.py:49-167): random, raster (TLBR) or highest-confidence order."""
        from . import _lib
        self.eval()
        input_ids = self._input_ids(sentences, max_text_length)
        self._check_device(input_ids)
        B, dev = input_ids.shape[0], input_ids.device
        grid_size, code_dim = 8, self.dims.feat_dim
        n_grids = grid_size ** 2
        if n_steps is None:
            n_steps = n_grids
        visual_pos = self._visual_pos(B, grid_size, dev)
        intermediate_imgs = []
        if position_random:
            positions = list(range(n_grids))
            rng = random.Random(seed) if seed is not None else random
            rng.shuffle(positions)
            if n_steps > n_grids:
                extra = list(range(n_steps - n_grids))
                (random.Random(seed) if seed is not None else random).shuffle(extra)
                positions = extra + positions
        lib = _lib.load()
        code = torch.empty(B, n_grids, code_dim, device=dev, dtype=torch.float32)
        vis_mask = torch.empty(B, n_grids, dtype=torch.uint8, device=dev)
        visited = torch.zeros(B, n_grids, dtype=torch.uint8, device=dev)
        self._nar_update(code, vis_mask, None, None, n_grids)          # every cell masked, code = mask_feat (:95-99)
        table, mf = self.vis_emb.weight, self._mask_feat32()
        pred_prob = pred_code_id = None
        lang = self.bert.language_stack(input_ids, input_ids > 0) if cache_language else None
        for i in range(n_steps):
            stream = torch.cuda.current_stream().cuda_stream
            position = -1                                              # confidence order: chosen on the device
            if position_random:
                position = positions.pop() % n_grids
                _lib.check("xlx_sampler_remask_cell",
                           lib.xlx_sampler_remask_cell(code.data_ptr(), vis_mask.data_ptr(), mf.data_ptr(), B, n_grids,
                                                       code_dim, position, stream))
            elif position_TLBR:
                position = i
            pred_prob, pred_code_id = self._predict(input_ids, code, visual_pos, lang)
            _lib.check("xlx_sampler_ar_update",
                       lib.xlx_sampler_ar_update(code.data_ptr(), vis_mask.data_ptr(), visited.data_ptr(),
                                                 pred_prob.data_ptr(), pred_code_id.data_ptr(), table.data_ptr(), B,
                                                 n_grids, code_dim, position, stream))
            if return_intermediate:
                intermediate_imgs.append(self._decode(code, B, code_dim, grid_size))
        if return_intermediate:
            return intermediate_imgs
        if return_codes:
            return code, pred_prob, pred_code_id
        return self._decode(code, B, code_dim, grid_size)
