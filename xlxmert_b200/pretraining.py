"""``XLxmertForPretraining`` (``x-lxmert/src/lxrt/modeling.py:56-308``) on the sm_100a library.

Same attribute names (``bert``, ``cls``, ``obj_predict_head``, ``mask_feat``, ``vis_emb``), same state-dict
keys, same ``forward`` keywords and the same output dict (``lm_loss`` / ``matched_loss`` / ``obj_loss`` /
``vis_loss`` detached, ``total_loss`` differentiable), so ``Trainer.forward`` (``lxmert_pretrain.py:143-225``)
and the sampler (``tasks/imggen_model.py``) call it unchanged.  Default = the canonical task set of
``pretrain.bash:24-26``: MaskLM + ObjPredict + Matched, ``--visualLosses obj``.  ``visual_losses=("obj", "feat")`` adds the
feature-regression loss (``modeling.py:270-284``; the published default ``--visualLosses obj,feat``, ``param.py:123``)
and ``task_qa=True`` the answer head whose loss joins every task's (``--taskQA``, ``modeling.py:89-90,286-299``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .config import LxmertDims
from .encoder import dims_from_hf_config
from .heads import B200LxmertPreTrainingHeads, B200LxmertVisualAnswerHead, B200LxmertVisualObjHead
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
    def __init__(self, config, num_clusters: int = 10000, passes: int = 3, visual_losses=None,
                 task_qa: Optional[bool] = None, num_qa_labels: Optional[int] = None):
        super().__init__()
        hf = not isinstance(config, LxmertDims)
        dims = dims_from_hf_config(config) if hf else config
        if dims.num_clusters != num_clusters:
            dims = LxmertDims(**{**dims.asdict(), "num_clusters": num_clusters})
        self.dims = dims
        self.config = config if hf else None
        if hf:      # modeling.py:66-70,114-137 with the `feat` entry keyed on its own flag
            if visual_losses is None:
                visual_losses = [k for k, on in (("obj", getattr(config, "visual_obj_loss", True)),
                                                 ("feat", getattr(config, "visual_feat_loss", False))) if on]
            task_qa = getattr(config, "task_qa", False) if task_qa is None else task_qa
            num_qa_labels = getattr(config, "num_qa_labels", 9500) if num_qa_labels is None else num_qa_labels
        visual_losses = ("obj",) if visual_losses is None else tuple(visual_losses)
        bad = set(visual_losses) - {"obj", "feat"}
        if bad or not visual_losses:
            raise ValueError(f"visual_losses must be a non-empty subset of ('obj', 'feat') in cluster mode, got "
                             f"{visual_losses!r} ('attr' needs the non-clustering head, modeling.py:33-36)")
        self.task_mask_lm = self.task_obj_predict = self.task_matched = True
        self.task_qa = bool(task_qa)
        self.num_qa_labels = int(num_qa_labels) if num_qa_labels is not None else 9500
        self.bert = B200LxmertModel(dims, passes=passes)
        self.cls = B200LxmertPreTrainingHeads(dims, self.bert.embeddings.word_embeddings.weight, passes=passes)
        self.obj_predict_head = B200LxmertVisualObjHead(dims, num_clusters, passes=passes)
        if self.task_qa:
            self.answer_head = B200LxmertVisualAnswerHead(dims, self.num_qa_labels, passes=passes)
        self.mask_feat = nn.Parameter(torch.zeros(dims.feat_dim))
        self.vis_emb: Optional[nn.Embedding] = None
        self.visual_losses = {}
        if "obj" in visual_losses:
            self.visual_losses["obj"] = {"shape": (-1,), "num": num_clusters, "loss": "visual_ce"}
        if "feat" in visual_losses:
            self.visual_losses["feat"] = {"shape": (-1, dims.feat_dim), "num": dims.feat_dim, "loss": "l2"}
        self.obj_predict_head.visual_losses = {k: {"shape": v["shape"], "num": v["num"]}
                                               for k, v in self.visual_losses.items()}

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
            keys = self.visual_losses
            vl = self.obj_predict_head.losses(
                visual_output, obj_labels=label_dict['obj_labels'] if 'obj' in keys else None,
                feat_labels=label_dict['feat_labels'] if 'feat' in keys else None, vis_mask=vis_mask)
            for key in keys:                                                  # modeling.py:242-284, same order
                total_loss = vl[key] if total_loss is None else total_loss + vl[key]
                out_dict[f'{key}_loss'] = vl[key].detach()
            out_dict['vis_loss'] = total_loss.detach()
        elif not (task == 'qa' and self.task_qa):
            raise ValueError(f"task must be one of 'word_mask', 'vis_mask', 'matched'"
                             f"{' or qa' if self.task_qa else ''} (got {task!r})")
        if self.task_qa:                                                      # modeling.py:286-299: on every task
            qa_loss, qa_pred = self.answer_head.loss(pooled_output, label_dict['qa_labels'])
            total_loss = qa_loss if total_loss is None else total_loss + qa_loss
            out_dict['qa_loss'] = qa_loss.detach()
            out_dict['qa_pred'] = qa_pred
        out_dict['total_loss'] = total_loss
        return out_dict
