"""ctypes binding of ``libxlxmert_b200.so`` (the C ABI in ``include/xlxmert_b200.h``).

There is no fallback: if the library has not been built, ``load()`` raises — the product path never
routes around the CUDA extension.
"""
from __future__ import annotations

import ctypes as C
import os

from .config import LxmertDims

_LIB = None


class XlxDims(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("heads", C.c_int32), ("intermediate", C.c_int32),
                ("feat_dim", C.c_int32), ("pos_dim", C.c_int32), ("l_layers", C.c_int32),
                ("r_layers", C.c_int32), ("x_layers", C.c_int32), ("ln_eps", C.c_float)]

    @classmethod
    def from_dims(cls, d: LxmertDims) -> "XlxDims":
        return cls(d.hidden, d.heads, d.intermediate, d.feat_dim, d.pos_dim, d.l_layers, d.r_layers,
                   d.x_layers, d.ln_eps)


class XlxDropout(C.Structure):
    """``xlx_dropout`` (include/xlxmert_b200.h): probabilities + the step's seed."""
    _fields_ = [("p_hidden", C.c_float), ("p_attn", C.c_float), ("seed", C.c_uint64)]


class XlxError(RuntimeError):
    def __init__(self, fn: str, code: int):
        self.code = code
        msg = load().xlx_strerror(code).decode()
        super().__init__(f"{fn} failed with code {code}: {msg}")


def step_dropout(module, dims):
    """``xlx_dropout`` for one training-mode forward of ``module`` (None when dropout is off): probabilities from the
    dims, a fresh 62-bit seed drawn on the HOST from torch's default CPU generator (reproducible under
    ``torch.manual_seed``, no device sync)."""
    import torch
    if not module.training or (dims.hidden_dropout <= 0.0 and dims.attention_dropout <= 0.0):
        return None
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    return XlxDropout(float(dims.hidden_dropout), float(dims.attention_dropout), seed)


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libxlxmert_b200.so")


def _sig(lib):
    P, I32, I64, SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    D = C.POINTER(XlxDims)
    lib.xlx_version.restype = C.c_char_p
    lib.xlx_strerror.restype = C.c_char_p
    lib.xlx_strerror.argtypes = [I32]
    lib.xlx_launch_count.restype = I64
    lib.xlx_gemm_launch_count.restype = I64
    lib.xlx_profile_gemm_begin.restype = None
    lib.xlx_profile_gemm_end.restype = I32
    lib.xlx_profile_gemm_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(I64)]
    for name in ("xlx_encoder_num_params", "xlx_encoder_grad_elems"):
        getattr(lib, name).restype = I64
        getattr(lib, name).argtypes = [D]
    for name in ("xlx_encoder_param_elems", "xlx_encoder_grad_offset"):
        getattr(lib, name).restype = I64
        getattr(lib, name).argtypes = [D, I64]
    lib.xlx_encoder_prep_bytes.restype = SZ
    lib.xlx_encoder_prep_bytes.argtypes = [D]
    lib.xlx_encoder_prepare.restype = I32
    lib.xlx_encoder_prepare.argtypes = [D, C.c_void_p, P, P]
    lib.xlx_encoder_workspace_bytes.restype = SZ
    lib.xlx_encoder_workspace_bytes.argtypes = [D, I32, I32, I32, I32]
    lib.xlx_encoder_fwd.restype = I32
    DR = C.POINTER(XlxDropout)
    lib.xlx_encoder_fwd.argtypes = [D, P, P, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, SZ, I32, I32, DR, P]
    lib.xlx_dropout_mask.restype = I32
    lib.xlx_dropout_mask.argtypes = [C.c_uint64, C.c_uint32, C.c_float, I32, I64, I32, P, P]
    lib.xlx_dropout_site.restype = I32
    lib.xlx_dropout_site.argtypes = [D, I32, I32, I32]
    lib.xlx_encoder_bwd.restype = I32
    lib.xlx_encoder_bwd.argtypes = [D, P, P, I32, I32, I32, P, P, P, P, P, P, P, SZ, I32, I32, DR, P, P]
    lib.xlx_encoder_probs_offset.restype = I64
    lib.xlx_encoder_probs_offset.argtypes = [D, I32, I32, I32, I32, I32]
    lib.xlx_encoder_grad_stage_range.restype = I32
    lib.xlx_encoder_grad_stage_range.argtypes = [D, I32, C.POINTER(I64), C.POINTER(I64)]
    PP = P
    lib.xlx_embeddings_save_bytes.restype = SZ
    lib.xlx_embeddings_save_bytes.argtypes = [D, I32, I32]
    lib.xlx_embeddings_scratch_bytes.restype = SZ
    lib.xlx_embeddings_scratch_bytes.argtypes = [D, I32, I32]
    lib.xlx_embeddings_fwd.restype = I32
    lib.xlx_embeddings_fwd.argtypes = [D, I32, I32, P, P, P, PP, P, P, DR, P]
    lib.xlx_embeddings_bwd.restype = I32
    lib.xlx_embeddings_bwd.argtypes = [D, I32, I32, I32, I32, I32, P, P, PP, P, P, PP, P, P, SZ, DR, P]
    lib.xlx_pooler_workspace_bytes.restype = SZ
    lib.xlx_pooler_workspace_bytes.argtypes = [D, I32]
    lib.xlx_pooler_fwd.restype = I32
    lib.xlx_pooler_fwd.argtypes = [D, I32, I32, P, P, P, P, P, SZ, I32, P]
    lib.xlx_pooler_bwd.restype = I32
    lib.xlx_pooler_bwd.argtypes = [D, I32, I32, P, P, P, P, P, P, SZ, I32, P]
    for kind in ("objhead", "lmhead", "qahead"):
        f = getattr(lib, f"xlx_{kind}_prep_bytes"); f.restype = SZ; f.argtypes = [D, I32]
        f = getattr(lib, f"xlx_{kind}_prepare"); f.restype = I32; f.argtypes = [D, I32, P, P, P]
        f = getattr(lib, f"xlx_{kind}_workspace_bytes"); f.restype = SZ; f.argtypes = [D, I32, I32]
    lib.xlx_objhead_fwd.restype = I32
    lib.xlx_objhead_fwd.argtypes = [D, I32, P, P, I32, P, P, P, P, P, P, P, P, P, P, P, SZ, I32, P]
    lib.xlx_objhead_bwd.restype = I32
    lib.xlx_objhead_bwd.argtypes = [D, I32, P, P, I32, P, P, P, P, P, P, P, P, P, P, SZ, I32, P]
    lib.xlx_feat_row_weight.restype = I32
    lib.xlx_feat_row_weight.argtypes = [P, I32, I32, I32, P, I32, P, P]
    for kind in ("lmhead", "qahead"):
        f = getattr(lib, f"xlx_{kind}_fwd"); f.restype = I32
        f.argtypes = [D, I32, P, P, I32, P, P, P, P, P, P, P, SZ, I32, P]
        f = getattr(lib, f"xlx_{kind}_bwd"); f.restype = I32
        f.argtypes = [D, I32, P, P, I32, P, P, P, P, P, P, SZ, I32, P]
    lib.xlx_matchhead_bwd_scores.restype = I32
    lib.xlx_matchhead_bwd_scores.argtypes = [D, I32, P, P, P, P, P, P, P]
    lib.xlx_visual_input_fwd.restype = I32
    lib.xlx_visual_input_fwd.argtypes = [P, P, P, P, I32, I32, P, P]
    lib.xlx_visual_input_bwd.restype = I32
    lib.xlx_visual_input_bwd.argtypes = [P, P, I32, I32, P, P, P]
    lib.xlx_generator_num_params.restype = I64
    lib.xlx_generator_launch_count.restype = I64
    lib.xlx_generator_prep_bytes.restype = SZ
    lib.xlx_generator_workspace_bytes.restype = SZ
    lib.xlx_generator_workspace_bytes.argtypes = [I32]
    lib.xlx_generator_prepare.restype = I32
    lib.xlx_generator_prepare.argtypes = [P, P, P]
    lib.xlx_generator_fwd.restype = I32
    lib.xlx_generator_fwd.argtypes = [P, P, I32, P, P, P, P, P, P, SZ, I32, P]
    F64 = C.c_double
    lib.xlx_optim_scratch_floats.restype = I64
    lib.xlx_optim_scratch_floats.argtypes = [P, I32]
    lib.xlx_optim_launch_count.restype = I64
    lib.xlx_grad_sqnorm.restype = I32
    lib.xlx_grad_sqnorm.argtypes = [P, P, I32, P, P, P]
    lib.xlx_adamw_step.restype = I32
    lib.xlx_adamw_step.argtypes = [P, P, P, P, P, P, I32, F64, F64, F64, F64, I32, I32, P, F64, P]
    lib.xlx_kmeans_prep_bytes.restype = SZ
    lib.xlx_kmeans_prep_bytes.argtypes = [I32, I32]
    lib.xlx_kmeans_prepare.restype = I32
    lib.xlx_kmeans_prepare.argtypes = [I32, I32, P, P, P]
    lib.xlx_kmeans_workspace_bytes.restype = SZ
    lib.xlx_kmeans_workspace_bytes.argtypes = [I32, I32, I32]
    lib.xlx_kmeans_assign.restype = I32
    lib.xlx_kmeans_assign.argtypes = [I32, I32, P, I32, P, P, P, P, SZ, I32, P]
    lib.xlx_labelled_rows.restype = I32
    lib.xlx_labelled_rows.argtypes = [P, I32, I64, P, P, P]
    lib.xlx_gather_rows.restype = I32
    lib.xlx_gather_rows.argtypes = [P, P, P, I32, I32, P, P, P]
    lib.xlx_scatter_rows.restype = I32
    lib.xlx_scatter_rows.argtypes = [P, P, I32, I32, I32, P, P]
    lib.xlx_pretrain_inputs_layout.restype = I32
    lib.xlx_pretrain_inputs_layout.argtypes = [I32, I32, I32, C.POINTER(I64), C.POINTER(I64)]
    lib.xlx_pretrain_inputs_unpack.restype = I32
    lib.xlx_pretrain_inputs_unpack.argtypes = [P, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P]
    lib.xlx_sampler_nar_update.restype = I32
    lib.xlx_sampler_nar_update.argtypes = [P, P, P, P, P, P, I32, I32, I32, I32, P, P]
    lib.xlx_sampler_ar_update.restype = I32
    lib.xlx_sampler_ar_update.argtypes = [P, P, P, P, P, P, I32, I32, I32, I32, P]
    lib.xlx_sampler_remask_cell.restype = I32
    lib.xlx_sampler_remask_cell.argtypes = [P, P, P, I32, I32, I32, I32, P]
    lib.xlx_matchhead_scratch_floats.restype = I64
    lib.xlx_matchhead_scratch_floats.argtypes = [I32]
    lib.xlx_matchhead_fwd.restype = I32
    lib.xlx_matchhead_fwd.argtypes = [D, I32, P, P, P, P, P, P, P, P]
    lib.xlx_matchhead_bwd.restype = I32
    lib.xlx_matchhead_bwd.argtypes = [D, I32, P, P, P, P, P, P, P, P, P, P]


def load():
    """The loaded library; raises if it has not been built (``python -m xlxmert_b200.build``)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: the B200 hot path has no CPU/PyTorch fallback. "
                "Build it with `python -m xlxmert_b200.build` (needs nvcc).")
        lib = C.CDLL(path)
        _sig(lib)
        _LIB = lib
    return _LIB


def check(fn: str, code: int) -> None:
    if code != 0:
        raise XlxError(fn, int(code))
