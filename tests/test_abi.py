"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/xlxmert_b200.h declares,
its size/introspection entry points answer without a GPU, and the product path refuses to run without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "xlxmert_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xlx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from xlxmert_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.xlx_version()
    assert lib.xlx_strerror(-22).decode().startswith("sequence length")


def test_size_queries_work_without_gpu():
    from xlxmert_b200 import _lib, params as P
    from xlxmert_b200.config import DEFAULT_DIMS as D
    lib = _lib.load()
    cd = _lib.XlxDims.from_dims(D)
    n = lib.xlx_encoder_num_params(C.byref(cd))
    assert n == P.num_encoder_params(D) == len(P.encoder_param_names(D))
    total = lib.xlx_encoder_grad_elems(C.byref(cd))
    assert total == sum(lib.xlx_encoder_param_elems(C.byref(cd), i) for i in range(n))
    specs = P.encoder_param_specs(D)
    for i in (0, 5, 8, 17, n - 1):
        shape = specs[i][1]
        assert lib.xlx_encoder_param_elems(C.byref(cd), i) == int(torch.tensor(shape).prod())
    assert lib.xlx_encoder_workspace_bytes(C.byref(cd), 256, 20, 64, 1) > lib.xlx_encoder_workspace_bytes(C.byref(cd), 256, 20, 64, 0) > 0
    assert lib.xlx_encoder_workspace_bytes(C.byref(cd), 2, 65, 64, 0) == 0          # unsupported length → 0, not a crash
    assert lib.xlx_objhead_workspace_bytes(C.byref(cd), 10000, 128) > 0
    assert lib.xlx_generator_num_params() == 150 and lib.xlx_generator_workspace_bytes(2) > 0
    bad = _lib.XlxDims.from_dims(D)
    bad.hidden = 100
    assert lib.xlx_encoder_num_params(C.byref(bad)) == -20


def test_generator_state_dict_contract():
    """State-dict keys of the drop-in generator are exactly the reference's."""
    from xlxmert_b200 import params as P
    from xlxmert_b200.generator import B200Generator
    G = B200Generator()
    want = set(P.init_generator_state_dict(seed=0).keys())
    assert set(G.state_dict().keys()) == want
    assert sum(1 for k in want if k.endswith("weight_orig")) == 17      # 17 spectrally normalised convs


def test_pretraining_state_dict_contract():
    from xlxmert_b200.config import TINY_DIMS
    from xlxmert_b200.pretraining import B200XLxmertForPretraining
    m = B200XLxmertForPretraining(TINY_DIMS, num_clusters=TINY_DIMS.num_clusters)
    m.set_visual_embedding(torch.rand(TINY_DIMS.num_clusters, TINY_DIMS.feat_dim))
    keys = set(m.state_dict().keys())
    for k in ("mask_feat", "bert.embeddings.word_embeddings.weight", "bert.encoder.visn_fc.visn_fc.weight",
              "bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.r_layers.0.output.LayerNorm.bias",
              "bert.encoder.x_layers.1.visual_attention.att.key.bias", "bert.encoder.x_layers.0.lang_inter.dense.weight",
              "bert.pooler.dense.weight", "cls.predictions.bias", "cls.predictions.transform.dense.weight",
              "cls.predictions.decoder.weight", "cls.seq_relationship.weight", "obj_predict_head.transform.LayerNorm.weight",
              "obj_predict_head.linear_feat.weight", "obj_predict_head.out_cluster.weight",
              "obj_predict_head.out_cluster.bias", "vis_emb.weight"):
        assert k in keys, k
    assert m.cls.predictions.decoder.weight is m.bert.embeddings.word_embeddings.weight       # modeling.py:86
    assert m.obj_predict_head.out_cluster.weight is m.vis_emb.weight                          # modeling.py:151
    assert not m.vis_emb.weight.requires_grad


def test_product_path_has_no_cpu_fallback():
    from xlxmert_b200.config import TINY_DIMS
    from xlxmert_b200.lxmert import B200LxmertModel
    m = B200LxmertModel(TINY_DIMS)
    ids = torch.ones(1, 4, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(input_ids=ids, visual_feats=torch.zeros(1, 4, TINY_DIMS.feat_dim), visual_pos=torch.zeros(1, 4, 4))


def test_backward_stage_ranges_tile_the_gradient_arena():
    """The four backward stages complete disjoint arena ranges that together cover it (what the overlapped
    data-parallel exchange relies on)."""
    from xlxmert_b200 import _lib
    from xlxmert_b200.config import DEFAULT_DIMS as D, TINY_DIMS
    lib = _lib.load()
    for dims in (D, TINY_DIMS):
        cd = _lib.XlxDims.from_dims(dims)
        total = lib.xlx_encoder_grad_elems(C.byref(cd))
        spans = []
        for stage in (1, 2, 4, 8):
            off, n = C.c_int64(), C.c_int64()
            assert lib.xlx_encoder_grad_stage_range(C.byref(cd), stage, C.byref(off), C.byref(n)) == 0
            spans.append((off.value, n.value))
        spans.sort()
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
        off, n = C.c_int64(), C.c_int64()
        assert lib.xlx_encoder_grad_stage_range(C.byref(cd), 3, C.byref(off), C.byref(n)) == -25


def test_stage_mask_ranges_and_host_side_argument_checks():
    """Host logic that needs no GPU: the module-level union of backward-stage ranges (what the data-parallel backward
    all-reduces after each call), the size queries of the late additions, and argument errors that must be raised
    before anything touches a device."""
    import numpy as np
    from xlxmert_b200 import _lib
    from xlxmert_b200.config import TINY_DIMS
    from xlxmert_b200.encoder import B200LxmertEncoder, _BWD_STAGES
    from xlxmert_b200.kmeans import B200IndexFlatL2
    lib = _lib.load()
    enc = B200LxmertEncoder(dims=TINY_DIMS)
    total = lib.xlx_encoder_grad_elems(C.byref(enc._cdims))
    assert _BWD_STAGES == (1, 14)
    (o1, n1), (o2, n2) = enc._stage_range(1), enc._stage_range(14)
    assert o2 == 0 and o2 + n2 == o1 and o1 + n1 == total            # [everything below | cross-modality layers]
    assert enc._stage_range(15) == (0, total)
    with pytest.raises(ValueError):
        enc._stage_range(1 | 4)                                       # cross + language: not adjacent in the arena
    # parameter list cache: same objects, rebuilt after a parameter is replaced inside a layer
    p0 = enc._param_list()
    assert enc._param_list() is p0
    lin = enc.layer[0].attention.self.query
    lin.weight = torch.nn.Parameter(lin.weight.detach().clone())
    p1 = enc._param_list()
    assert p1 is not p0 and any(a is lin.weight for a in p1)
    # k-means index: faiss call shape, loud failures
    assert lib.xlx_kmeans_prep_bytes(2048, 10000) > 2 * 10000 * 2048 * 2
    assert lib.xlx_kmeans_prep_bytes(2047, 10000) == 0 and lib.xlx_kmeans_workspace_bytes(2048, 10000, 0) == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200IndexFlatL2(8, device="cpu").add(np.zeros((4, 8), np.float32))
    with pytest.raises(ValueError):
        B200IndexFlatL2(8).add(np.zeros((4, 9), np.float32))
    with pytest.raises(RuntimeError, match="before add"):
        B200IndexFlatL2(8).search(np.zeros((1, 8), np.float32), 1)
    with pytest.raises(NotImplementedError):
        B200IndexFlatL2(8).search(np.zeros((1, 8), np.float32), 5)


def test_ctypes_signatures_match_the_header():
    """Every entry point `_lib.py` declares argtypes for takes exactly as many parameters as `include/xlxmert_b200.h`
    says — a changed C signature with a stale ctypes list would pass garbage, not fail."""
    import __graft_entry__ as entry
    entry.build()
    from xlxmert_b200 import _lib
    lib = _lib.load()
    src = open(os.path.join(ROOT, "include", "xlxmert_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    decls = re.findall(r"\b(xlx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(decls) >= 60
    checked = 0
    for name, params in decls:
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
        checked += 1
    assert checked >= 40, checked


def test_round2_entry_points_validate_arguments_without_gpu():
    """Size queries and argument checks of the answer head, the feature-regression weights and the packed inputs run on
    the host: they must answer (or refuse with the documented codes) before anything touches the device."""
    from xlxmert_b200 import _lib
    from xlxmert_b200.config import DEFAULT_DIMS as D
    lib = _lib.load()
    cd = _lib.XlxDims.from_dims(D)
    assert lib.xlx_qahead_prep_bytes(C.byref(cd), 9500) > lib.xlx_lmhead_prep_bytes(C.byref(cd), 9500)      # 2H-wide transform
    assert lib.xlx_qahead_workspace_bytes(C.byref(cd), 9500, 256) > 0
    assert lib.xlx_qahead_workspace_bytes(C.byref(cd), 9500, 0) == 0
    bad = _lib.XlxDims.from_dims(D)
    bad.hidden = 100
    assert lib.xlx_qahead_prep_bytes(C.byref(bad), 9500) == 0
    assert lib.xlx_feat_row_weight(None, 0, 64, 2048, None, 0, None, None) == -21       # B < 1
    assert lib.xlx_feat_row_weight(None, 2, 64, 2048, None, 5, None, None) == -21       # rows NULL needs n == B·V
    assert lib.xlx_feat_row_weight(None, 2, 64, 2048, None, 128, None, None) == -24     # NULL pointers
    offs, total = (C.c_int64 * 7)(), C.c_int64()
    assert lib.xlx_pretrain_inputs_layout(256, 20, 64, offs, C.byref(total)) == 0
    assert list(offs) == sorted(offs) and offs[6] + 256 * 8 <= total.value == 495616
    assert lib.xlx_matchhead_bwd_scores(C.byref(cd), 0, None, None, None, None, None, None, None) == -21
