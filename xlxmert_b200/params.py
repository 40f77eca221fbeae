"""Parameter tables: names, shapes and the canonical order shared by Python and the C-ABI.

The names are the reference's state-dict keys (SURVEY.md §8b "State-dict contract"): HF
``LxmertEncoder`` (``modeling_lxmert.py:487-504``) holds ``visn_fc``, ``layer`` (language),
``r_layers`` (vision) and ``x_layers`` (cross-modality).  ``include/xlxmert_b200.h`` documents the
same order as ``XLX_P_*`` slot indices; ``encoder_param_names`` below is its Python twin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .config import LxmertDims

# ---- per-block slot lists (order matters: it is the C-ABI order) -------------------------------

ATT_SLOTS = [  # LxmertAttention + LxmertAttentionOutput (HF:217-288); q/k/v weights (and biases) are
    # adjacent so that the fused-QKV weight gradient [3H, H] lands on three consecutive slots
    ("{a}.query.weight", "HH"), ("{a}.key.weight", "HH"), ("{a}.value.weight", "HH"),
    ("{a}.query.bias", "H"), ("{a}.key.bias", "H"), ("{a}.value.bias", "H"),
    ("{o}.dense.weight", "HH"), ("{o}.dense.bias", "H"),
    ("{o}.LayerNorm.weight", "H"), ("{o}.LayerNorm.bias", "H"),
]
# Order in which ``init_state_dict`` draws the attention parameters.  The committed goldens
# (tests/golden/*.npz) were generated with this order; it only fixes RNG consumption, not the ABI.
ATT_INIT_ORDER = [
    ("{a}.query.weight", "HH"), ("{a}.query.bias", "H"),
    ("{a}.key.weight", "HH"), ("{a}.key.bias", "H"),
    ("{a}.value.weight", "HH"), ("{a}.value.bias", "H"),
    ("{o}.dense.weight", "HH"), ("{o}.dense.bias", "H"),
    ("{o}.LayerNorm.weight", "H"), ("{o}.LayerNorm.bias", "H"),
]
FFN_SLOTS = [  # LxmertIntermediate + LxmertOutput (HF:327-350)
    ("{i}.dense.weight", "IH"), ("{i}.dense.bias", "I"),
    ("{o}.dense.weight", "HI"), ("{o}.dense.bias", "H"),
    ("{o}.LayerNorm.weight", "H"), ("{o}.LayerNorm.bias", "H"),
]
VISN_SLOTS = [  # LxmertVisualFeatureEncoder (HF:460-484)
    ("visn_fc.visn_fc.weight", "HF"), ("visn_fc.visn_fc.bias", "H"),
    ("visn_fc.visn_layer_norm.weight", "H"), ("visn_fc.visn_layer_norm.bias", "H"),
    ("visn_fc.box_fc.weight", "HP"), ("visn_fc.box_fc.bias", "H"),
    ("visn_fc.box_layer_norm.weight", "H"), ("visn_fc.box_layer_norm.bias", "H"),
]

N_ATT = len(ATT_SLOTS)      # 10
N_FFN = len(FFN_SLOTS)      # 6
N_VISN = len(VISN_SLOTS)    # 8
N_LAYER = N_ATT + N_FFN     # 16: one LxmertLayer
N_XLAYER = 3 * N_ATT + 2 * N_FFN  # 42: cross, lang self, visn self, lang ffn, visn ffn


def _shape(code: str, d: LxmertDims) -> Tuple[int, ...]:
    m = {"H": d.hidden, "I": d.intermediate, "F": d.feat_dim, "P": d.pos_dim}
    return tuple(m[c] for c in code)


def _att(prefix_att: str, prefix_out: str, init_order: bool = False):
    return [(n.format(a=prefix_att, o=prefix_out), s) for n, s in (ATT_INIT_ORDER if init_order else ATT_SLOTS)]


def _ffn(prefix_inter: str, prefix_out: str):
    return [(n.format(i=prefix_inter, o=prefix_out), s) for n, s in FFN_SLOTS]


def encoder_param_specs(d: LxmertDims, init_order: bool = False) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every ``LxmertEncoder`` parameter in C-ABI slot order (``init_order=True``: the
    RNG-draw order of the seeded test weights, see ``ATT_INIT_ORDER``)."""
    from functools import partial
    _att = partial(globals()["_att"], init_order=init_order)
    out = list(VISN_SLOTS)
    for i in range(d.l_layers):
        p = f"layer.{i}"
        out += _att(f"{p}.attention.self", f"{p}.attention.output")
        out += _ffn(f"{p}.intermediate", f"{p}.output")
    for i in range(d.r_layers):
        p = f"r_layers.{i}"
        out += _att(f"{p}.attention.self", f"{p}.attention.output")
        out += _ffn(f"{p}.intermediate", f"{p}.output")
    for i in range(d.x_layers):
        p = f"x_layers.{i}"
        out += _att(f"{p}.visual_attention.att", f"{p}.visual_attention.output")
        out += _att(f"{p}.lang_self_att.self", f"{p}.lang_self_att.output")
        out += _att(f"{p}.visn_self_att.self", f"{p}.visn_self_att.output")
        out += _ffn(f"{p}.lang_inter", f"{p}.lang_output")
        out += _ffn(f"{p}.visn_inter", f"{p}.visn_output")
    return [(n, _shape(s, d)) for n, s in out]


def encoder_param_names(d: LxmertDims) -> List[str]:
    return [n for n, _ in encoder_param_specs(d)]


def num_encoder_params(d: LxmertDims) -> int:
    return N_VISN + (d.l_layers + d.r_layers) * N_LAYER + d.x_layers * N_XLAYER


def model_param_specs(d: LxmertDims) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every ``LxmertModel`` parameter: embeddings + encoder + pooler."""
    H = d.hidden
    emb = [
        ("embeddings.word_embeddings.weight", (d.vocab, H)),
        ("embeddings.position_embeddings.weight", (d.max_pos, H)),
        ("embeddings.token_type_embeddings.weight", (d.type_vocab, H)),
        ("embeddings.LayerNorm.weight", (H,)), ("embeddings.LayerNorm.bias", (H,)),
    ]
    enc = [("encoder." + n, s) for n, s in encoder_param_specs(d, init_order=True)]
    pool = [("pooler.dense.weight", (H, H)), ("pooler.dense.bias", (H,))]
    return emb + enc + pool


def objhead_param_specs(d: LxmertDims) -> List[Tuple[str, Tuple[int, ...]]]:
    """``lxrt.modeling.LxmertVisualObjHead`` (x-lxmert/src/lxrt/modeling.py:8-53), cluster mode."""
    H, F, C = d.hidden, d.feat_dim, d.num_clusters
    return [
        ("transform.dense.weight", (H, H)), ("transform.dense.bias", (H,)),
        ("transform.LayerNorm.weight", (H,)), ("transform.LayerNorm.bias", (H,)),
        ("linear_feat.weight", (F, H)), ("linear_feat.bias", (F,)),
        ("out_cluster.weight", (C, F)), ("out_cluster.bias", (C,)),
    ]


def answerhead_param_specs(d: LxmertDims, num_answers: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """HF ``LxmertVisualAnswerHead`` (HF modeling_lxmert.py:610-623) — the reference's ``answer_head``
    (x-lxmert/src/lxrt/modeling.py:89-90)."""
    H = d.hidden
    return [("logit_fc.0.weight", (2 * H, H)), ("logit_fc.0.bias", (2 * H,)),
            ("logit_fc.2.weight", (2 * H,)), ("logit_fc.2.bias", (2 * H,)),
            ("logit_fc.3.weight", (num_answers, 2 * H)), ("logit_fc.3.bias", (num_answers,))]


def init_state_dict(specs, seed: int = 0, std: float = 0.02, dtype=torch.float32,
                    randomize_ln_bias: bool = False) -> Dict[str, torch.Tensor]:
    """Random initialisation in the spirit of HF ``_init_weights`` (``modeling_lxmert.py:668-680``):
    Linear/Embedding weights ~ N(0, 0.02), biases 0, LayerNorm weight 1 / bias 0, and rows
    ``padding_idx=0`` of the three embedding tables zero (``modeling_lxmert.py:184-186``).

    ``randomize_ln_bias=True`` perturbs biases and LayerNorm affine parameters as well, so that
    parity tests exercise every parameter (an all-zero bias hides indexing bugs).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in specs:
        if name.endswith("LayerNorm.weight") or name.endswith("layer_norm.weight"):
            t = torch.ones(shape, dtype=dtype)
            if randomize_ln_bias:
                t = t + 0.1 * torch.randn(shape, generator=g, dtype=dtype)
        elif name.endswith("bias"):
            t = torch.zeros(shape, dtype=dtype)
            if randomize_ln_bias:
                t = 0.05 * torch.randn(shape, generator=g, dtype=dtype)
        else:
            t = std * torch.randn(shape, generator=g, dtype=dtype)
            if "embeddings.weight" in name or name.endswith("_embeddings.weight"):
                t[0].zero_()
        sd[name] = t
    return sd


def count_params(specs) -> int:
    return sum(math.prod(s) for _, s in specs)


# ---- generator ----------------------------------------------------------------------------------

def generator_conv_specs(base_dim: int = 32, emb_dim: int = 2048, codebook_dim: int = 256,
                         n_blocks: int = 5, spade_hidden: int = 128):
    """Every convolution of ``image_generator/src/layers.py:Generator`` (:135-221) as
    ``(state-dict prefix, out_ch, in_ch_per_group, k, groups, spectral_norm)``, in forward order.
    Canonical run: base_dim 32, target 256 → 5 up-sampling blocks (train_generator.bash:6-8)."""
    c = base_dim
    convs = [("bottleneck_emb.0", codebook_dim, emb_dim, 1, 1, False),
             ("learned_init_conv.0", c, codebook_dim // 4, 3, 4, True),
             ("style_init_conv.0", c, codebook_dim // 4, 3, 4, True)]
    for i in range(n_blocks):
        p = f"resblocks.{i}"
        for cbn in ("cbn1", "cbn2"):
            convs += [(f"{p}.{cbn}.shared.0", spade_hidden, c, 3, 1, False),
                      (f"{p}.{cbn}.gamma", c, spade_hidden, 3, 1, False),
                      (f"{p}.{cbn}.beta", c, spade_hidden, 3, 1, False)]
        convs += [(f"{p}.conv1", c, c, 3, 1, True), (f"{p}.conv2", c, c, 3, 1, True),
                  (f"{p}.res_branch.1", c, c, 1, 1, True)]
    for i in range(n_blocks):
        convs.append((f"to_RGB_blocks.{i}.conv", 3, c, 3, 1, False))
    return convs


def init_generator_state_dict(seed: int = 0, power_iters: int = 50, **kw) -> Dict[str, torch.Tensor]:
    """Synthetic generator weights with *converged* spectral-norm buffers.

    Orthogonal conv weights and zero biases as ``Generator.init_parameter`` (layers.py:255-260);
    ``weight_u`` / ``weight_v`` are the leading singular pair of ``weight_orig`` (what ≥ 30 training
    forwards converge to — SURVEY §8d explains why un-converged buffers make parity vacuous).
    Biases get a small perturbation so that every parameter matters in parity tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    n_blocks = kw.get("n_blocks", 5)
    for prefix, co, ci, k, groups, sn in generator_conv_specs(**kw):
        w = torch.empty(co, ci, k, k)
        torch.nn.init.orthogonal_(w, generator=g)
        b = 0.02 * torch.randn(co, generator=g)
        if sn:
            wm = w.reshape(co, -1)
            u = torch.nn.functional.normalize(torch.randn(co, generator=g), dim=0)
            for _ in range(power_iters):
                v = torch.nn.functional.normalize(wm.t() @ u, dim=0, eps=1e-12)
                u = torch.nn.functional.normalize(wm @ v, dim=0, eps=1e-12)
            # keep sigma away from 1 so the division is visible in parity tests
            w = w * (1.0 + 0.25 * torch.rand((), generator=g))
            sd[prefix + ".bias"] = b
            sd[prefix + ".weight_orig"] = w
            sd[prefix + ".weight_u"] = u
            sd[prefix + ".weight_v"] = v
        else:
            sd[prefix + ".weight"] = w
            sd[prefix + ".bias"] = b
    for i in range(n_blocks):
        sd[f"resblocks.{i}.noise1.weight"] = torch.zeros(1)
        sd[f"resblocks.{i}.noise2.weight"] = torch.zeros(1)
    return sd
