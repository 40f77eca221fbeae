"""``XLxmertForPretraining`` (``x-lxmert/src/lxrt/modeling.py:56-308``) on the sm_100a library.

Same attribute names (``bert``, ``cls``, ``obj_predict_head``, ``mask_feat``, ``vis_emb``), same state-dict
keys, same ``forward`` keywords and the same output dict (``lm_loss`` / ``matched_loss`` / ``obj_loss`` /
``vis_loss`` detached, ``total_loss`` differentiable), so ``Trainer.forward`` (``lxmert_pretrain.py:143-225``)
and the sampler (``tasks/imggen_model.py``) call it unchanged.  Canonical task set of ``pretrain.bash:24-26``:
MaskLM + ObjPredict + Matched, ``--visualLosses obj``; the QA head is not part of that run and is not built.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .config import LxmertDims
from .encoder import dims_from_hf_config
from .heads import B200LxmertPreTrainingHeads, B200LxmertVisualObjHead
from .lxmert import B200LxmertModel


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _VisualInputFn(torch.autograd.Function):
    """``where(vis_mask, mask_feat, vis_emb(cluster_ids))`` (modeling.py:185-193)."""

    @staticmethod
    def forward(ctx, cluster_ids, vis_mask, table, mask_feat):
        lib = _lib.load()
        if not table.is_cuda:
            raise RuntimeError("the visual input runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, V = cluster_ids.shape
        F = table.shape[1]
        ids = cluster_ids.contiguous()
        m = None if vis_mask is None else vis_mask.reshape(B, V).to(torch.bool).contiguous()
        out = torch.empty(B, V, F, device=table.device, dtype=torch.float32)
        rc = lib.xlx_visual_input_fwd(table.data_ptr(), ids.data_ptr(), None if m is None else m.data_ptr(),
                                      mask_feat.data_ptr(), B * V, F, out.data_ptr(), _stream())
        _lib.check("xlx_visual_input_fwd", rc)
        ctx.m, ctx.shape = m, (B, V, F)
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.m is None:
            return None, None, None, None
        lib = _lib.load()
        B, V, F = ctx.shape
        d_out = d_out.contiguous().float()
        g = torch.empty(F, device=d_out.device, dtype=torch.float32)
        scratch = torch.empty(128 * F, device=d_out.device, dtype=torch.float32)
        rc = lib.xlx_visual_input_bwd(d_out.data_ptr(), ctx.m.data_ptr(), B * V, F, g.data_ptr(), scratch.data_ptr(),
                                      _stream())
        _lib.check("xlx_visual_input_bwd", rc)
        return None, None, None, g


class B200XLxmertForPretraining(nn.Module):
    def __init__(self, config, num_clusters: int = 10000, passes: int = 3):
        super().__init__()
        dims = config if isinstance(config, LxmertDims) else dims_from_hf_config(config)
        if dims.num_clusters != num_clusters:
            dims = LxmertDims(**{**dims.asdict(), "num_clusters": num_clusters})
        self.dims = dims
        self.config = None if isinstance(config, LxmertDims) else config
        self.task_mask_lm = self.task_obj_predict = self.task_matched = True
        self.task_qa = False
        self.bert = B200LxmertModel(dims, passes=passes)
        self.cls = B200LxmertPreTrainingHeads(dims, self.bert.embeddings.word_embeddings.weight, passes=passes)
        self.obj_predict_head = B200LxmertVisualObjHead(dims, num_clusters, passes=passes)
        self.mask_feat = nn.Parameter(torch.zeros(dims.feat_dim))
        self.vis_emb: Optional[nn.Embedding] = None
        self.visual_losses = {"obj": {"shape": (-1,), "num": num_clusters, "loss": "visual_ce"}}

    def set_visual_embedding(self, centroids):
        """modeling.py:140-151: frozen centroid table, tied to ``obj_predict_head.out_cluster.weight``."""
        if isinstance(centroids, np.ndarray):
            centroids = torch.from_numpy(centroids)
        centroids = centroids.to(device=self.mask_feat.device, dtype=torch.float32).contiguous()
        self.vis_emb = nn.Embedding.from_pretrained(centroids, freeze=True)
        self.obj_predict_head.out_cluster.weight = self.vis_emb.weight

    def visual_input(self, cluster_ids, vis_mask=None):
        return _VisualInputFn.apply(cluster_ids, vis_mask, self.vis_emb.weight, self.mask_feat)

    def forward(self, input_ids=None, visual_feats=None, visual_pos=None, attention_mask=None,
                visual_attention_mask=None, cluster_ids=None, vis_mask=None, token_type_ids=None, inputs_embeds=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, label_dict=None,
                task=('word_mask', 'vis_mask', 'matched', 'qa'), **kwargs):
        out_dict = {}
        if cluster_ids is not None:                                           # config.clustering (modeling.py:185)
            if self.vis_emb is None:
                raise RuntimeError("call set_visual_embedding(centroids) first")
            visual_feats = self.visual_input(cluster_ids, vis_mask if task == 'vis_mask' else None)
        elif task == 'vis_mask':
            visual_feats = torch.where(vis_mask.view(*visual_feats.shape[:2], 1).bool(),
                                       self.mask_feat.view(1, 1, -1).to(visual_feats.dtype), visual_feats)
        out = self.bert(input_ids=input_ids, visual_feats=visual_feats, visual_pos=visual_pos,
                        token_type_ids=token_type_ids, attention_mask=attention_mask,
                        visual_attention_mask=visual_attention_mask, inputs_embeds=inputs_embeds,
                        output_hidden_states=output_hidden_states, output_attentions=output_attentions)
        lang_output, visual_output, pooled_output = out[0], out[1], out[2]
        total_loss = None
        if task == 'word_mask':
            total_loss = self.cls.lm_loss(lang_output, label_dict['word_labels'])
            out_dict['lm_loss'] = total_loss.detach()
        elif task == 'matched':
            total_loss = self.cls.matched_loss(pooled_output, label_dict['matched_labels'])
            out_dict['matched_loss'] = total_loss.detach()
        elif task == 'vis_mask':
            total_loss = self.obj_predict_head.loss(visual_output, label_dict['obj_labels'])
            out_dict['obj_loss'] = total_loss.detach()
            out_dict['vis_loss'] = total_loss.detach()
        else:
            raise ValueError(f"task must be one of 'word_mask', 'vis_mask', 'matched' (got {task!r})")
        out_dict['total_loss'] = total_loss
        return out_dict
