"""ORACLE tooling — imports the *unmodified* reference classes in the build container.

``/root/reference`` is read-only and exists only in the build container (never on the GPU box), so
this module is used solely by ``oracle/make_golden.py`` and by CPU tests that skip when the
reference checkout is absent.  Nothing is copied from the reference; the shims below are
harness-side monkey patches that step around defects of the published code and the 4.1.1 → 5.5.0
``transformers`` API drift (§8c S1-S5).
"""
from __future__ import annotations

import importlib
import os
import sys

REF = os.environ.get("XLX_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "x-lxmert", "src", "lxrt"))


def import_lxrt_modeling():
    """``x-lxmert/src/lxrt/modeling.py`` with shims S1 (4.1.1 heads signature) and S2 (post_init)."""
    import torch  # noqa: F401
    import transformers.models.lxmert.modeling_lxmert as hf

    src = os.path.join(REF, "x-lxmert", "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    modeling = importlib.import_module("lxrt.modeling")

    class HeadsWithTiedDecoder(hf.LxmertPreTrainingHeads):  # S1 (V1)
        def __init__(self, config, embedding_weights=None):
            super().__init__(config)
            if embedding_weights is not None:
                self.predictions.decoder.weight = embedding_weights

    modeling.LxmertPreTrainingHeads = HeadsWithTiedDecoder

    cls = modeling.XLxmertForPretraining
    if not getattr(cls, "_xlx_shimmed", False):
        orig_init_weights = cls.init_weights

        def init_weights(self):  # S2 (V2)
            if not hasattr(self, "all_tied_weights_keys"):
                try:
                    self.post_init()
                    return
                except Exception:
                    pass
            orig_init_weights(self)

        cls.init_weights = init_weights
        cls._xlx_shimmed = True
    return modeling


def build_pretraining_model(num_clusters: int = 10000, keep_feat_loss: bool = False, **config_kw):
    """Reference ``XLxmertForPretraining`` with S3/S4 applied (obj loss only, ``--visualLosses obj``).
    ``keep_feat_loss``: leave the model as published — ``visual_losses`` = obj + feat (the default
    ``--visualLosses obj,feat`` of ``param.py:123``), so the vis_mask forward needs ``label_dict['feat_labels']``."""
    from transformers import LxmertConfig

    modeling = import_lxrt_modeling()
    cfg = LxmertConfig(**config_kw)
    cfg.num_clusters = num_clusters
    model = modeling.XLxmertForPretraining(cfg, num_clusters=num_clusters)
    model.config.n_centroids = num_clusters                      # S3 (D4)
    if not keep_feat_loss:
        model.visual_losses = {"obj": model.visual_losses["obj"]}    # S4 (D5)
        model.obj_predict_head.visual_losses = {"obj": model.obj_predict_head.visual_losses["obj"]}
    return model


def import_generator_layers():
    """``image_generator/src/layers.py`` (imports torchvision; present in this image)."""
    src = os.path.join(REF, "image_generator", "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    return importlib.import_module("layers")
