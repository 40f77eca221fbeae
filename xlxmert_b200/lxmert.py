"""Drop-in for HF ``LxmertModel`` (``transformers/models/lxmert/modeling_lxmert.py:683-830``) — the object the
reference holds as ``self.bert`` (``x-lxmert/src/lxrt/modeling.py:80``) and calls at ``modeling.py:195-206``,
``tasks/imggen_model.py:221-227`` and in every fine-tune model (``tasks/vqa_model.py:16``).

Same sub-module names (``embeddings``, ``encoder``, ``pooler``), same parameter names and shapes, same
``forward`` keyword arguments, and an output that indexes like HF's (``[0]`` language, ``[1]`` vision, ``[2]``
pooled).  The arithmetic runs in ``libxlxmert_b200.so``; there is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .config import LxmertDims
from .embeddings import B200LxmertEmbeddings, B200LxmertPooler
from .encoder import B200LxmertEncoder


class LxmertOutput(tuple):
    """Tuple ``(language_output, vision_output, pooled_output)`` that also answers to HF's attribute names
    (``LxmertModelOutput``, HF:60-90)."""

    def __new__(cls, lang, vis, pooled, lang_states=None, vis_states=None):
        self = super().__new__(cls, (lang, vis, pooled))
        self.language_output, self.vision_output, self.pooled_output = lang, vis, pooled
        self.language_hidden_states, self.vision_hidden_states = lang_states, vis_states
        self.language_attentions = self.vision_attentions = self.cross_encoder_attentions = None
        return self

    def with_attentions(self, lang_att, vis_att, cross_att):
        self.language_attentions, self.vision_attentions, self.cross_encoder_attentions = lang_att, vis_att, cross_att
        return self


class B200LxmertModel(nn.Module):
    def __init__(self, dims: LxmertDims, passes: int = 3, source: Optional[nn.Module] = None):
        super().__init__()
        self.dims = dims
        if source is not None:     # adopt the parameters of an existing HF LxmertModel
            self.embeddings = B200LxmertEmbeddings(dims, source=source.embeddings)
            self.encoder = B200LxmertEncoder(source.encoder, dims=dims, passes=passes)
            self.pooler = B200LxmertPooler(dims, passes=passes, source=source.pooler)
        else:
            self.embeddings = B200LxmertEmbeddings(dims)
            self.encoder = B200LxmertEncoder(dims=dims, passes=passes)
            self.pooler = B200LxmertPooler(dims, passes=passes)
        self.config = getattr(source, "config", None)

    @staticmethod
    def _additive(mask, B, S):
        """``(1 − mask)·finfo.min`` as ``[B,1,1,S]`` (HF:766-782)."""
        fmin = torch.finfo(torch.float32).min
        return ((1.0 - mask.to(torch.float32)) * fmin).view(B, 1, 1, S)

    @torch.no_grad()
    def language_stack(self, input_ids, attention_mask=None, token_type_ids=None):
        """Embeddings + the language-only layers: everything of the forward that depends on the text alone.  Pass the
        result as ``forward(..., language_stack=...)`` when the same sentences are encoded repeatedly against
        changing visual inputs (the sampling loops)."""
        B, L = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(B, L, dtype=torch.bool, device=input_ids.device)
        emb = self.embeddings(input_ids, token_type_ids)
        return self.encoder.language_stack(emb, self._additive(attention_mask, B, L))

    def forward(self, input_ids=None, visual_feats=None, visual_pos=None, attention_mask=None,
                visual_attention_mask=None, token_type_ids=None, inputs_embeds=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, language_stack=None, **kw):
        if visual_feats is None or visual_pos is None:
            raise ValueError("`visual_feats` and `visual_pos` cannot be `None`")           # HF:746-749
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")   # HF:737-738
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")      # HF:741-742
        ref = input_ids if input_ids is not None else inputs_embeds
        B, L = ref.shape[:2]
        if attention_mask is None:
            attention_mask = torch.ones(B, L, dtype=torch.bool, device=ref.device)         # HF:751-752
        # additive masks (HF:766-782): (1 − mask)·finfo.min, one row per sample
        fmin = torch.finfo(torch.float32).min
        prebuilt = getattr(attention_mask, "_xlx_additive", None)      # B200PretrainInputs built it in its unpack kernel
        if prebuilt is not None and prebuilt.shape == (B, L):
            lmask = prebuilt.view(B, 1, 1, L)
        else:
            lmask = ((1.0 - attention_mask.to(torch.float32)) * fmin).view(B, 1, 1, L)
        vmask = None
        if visual_attention_mask is not None:
            V = visual_feats.shape[1]
            vmask = ((1.0 - visual_attention_mask.to(torch.float32)) * fmin).view(B, 1, 1, V)
        want_hidden = bool(output_hidden_states)
        vis_att = lang_att = cross_att = None
        if language_stack is not None:
            if want_hidden or output_attentions:
                raise NotImplementedError("output_hidden_states / output_attentions with a cached language stack")
            (vis_states, _), (lang_states, _), _ = self.encoder(None, lmask, visual_feats, visual_pos, vmask,
                                                                 language_stack=language_stack)
        else:
            emb = self.embeddings(input_ids, token_type_ids, inputs_embeds)
            prev = self.encoder.output_hidden_states
            self.encoder.output_hidden_states = want_hidden
            try:
                (vis_states, vis_att), (lang_states, lang_att), cross_att = self.encoder(
                    emb, lmask, visual_feats, visual_pos, vmask, output_attentions=output_attentions)
            finally:
                self.encoder.output_hidden_states = prev
        lang, vis = lang_states[-1], vis_states[-1]
        pooled = self.pooler(lang)
        out = LxmertOutput(lang, vis, pooled, lang_states if want_hidden else None,
                           vis_states if want_hidden else None)
        if output_attentions:
            out.with_attentions(lang_att, vis_att, cross_att)
        return out
