"""GPU edge cases the reference's callers produce: batch 1 and odd batches, one-token text, one
visual position, the last partial batch, odd generator batches (an image pair straddling a 128-pixel MMA tile), a
non-trivial visual attention mask, and large-batch linearity properties checked without the oracle."""
import pytest
import torch

from oracle import generator_oracle as GO
from oracle import lxrt_oracle as O
from xlxmert_b200 import params as P
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D, TINY_DIMS

from util import rel_err

pytestmark = pytest.mark.gpu


def _enc(d, sd, **kw):
    from xlxmert_b200.encoder import B200LxmertEncoder
    e = B200LxmertEncoder(dims=d, **kw)
    e.load_state_dict(O.sub(sd, "encoder"), strict=True)
    return e.cuda()


@pytest.mark.parametrize("B,L,V", [(1, 1, 1), (1, 20, 64), (7, 3, 5), (2, 64, 64)])
def test_encoder_small_and_extreme_shapes(B, L, V):
    d = TINY_DIMS
    sd = P.init_state_dict(P.model_param_specs(d), seed=11, randomize_ln_bias=True)
    g = torch.Generator().manual_seed(B * 100 + L)
    emb = torch.randn(B, L, d.hidden, generator=g)
    feats = torch.randn(B, V, d.feat_dim, generator=g).abs()
    pos = torch.rand(B, V, 4, generator=g)
    am = torch.ones(B, L, dtype=torch.bool)
    if L > 2:
        am[0, L - 1] = False
    mask = O.extended_mask(am, torch.float32)
    vm = torch.ones(B, V, dtype=torch.bool)
    if V > 2:
        vm[-1, 0] = False
    vmask = O.extended_mask(vm, torch.float32)
    sdo = {k: v.clone().requires_grad_(True) for k, v in O.sub(sd, "encoder").items()}
    emb_o = emb.clone().requires_grad_(True)
    ls, vs = O.encoder(sdo, emb_o, mask, feats, pos, vmask, heads=d.heads, n_l=d.l_layers, n_r=d.r_layers, n_x=d.x_layers)
    (ls[-1].sum() + (vs[-1] ** 2).sum()).backward()
    enc = _enc(d, sd).train()
    emb_g = emb.cuda().requires_grad_(True)
    (v, _), (l, _), _ = enc(emb_g, mask.cuda(), feats.cuda(), pos.cuda(), vmask.cuda())
    assert rel_err(l[-1].detach().cpu(), ls[-1].detach()) < 1e-4
    assert rel_err(v[-1].detach().cpu(), vs[-1].detach()) < 1e-4
    (l[-1].sum() + (v[-1] ** 2).sum()).backward()
    assert rel_err(emb_g.grad.cpu(), emb_o.grad) < 1e-3
    name = "x_layers.0.visual_attention.att.value.weight"
    assert rel_err(dict(enc.named_parameters())[name].grad.cpu(), sdo[name].grad) < 1e-3


def test_encoder_batch_independence_at_full_size():
    """Samples are independent in every op of the path: the first rows of a B=256 forward equal a B=3
    forward of the same samples bit for bit (same kernels, same tile-local arithmetic)."""
    sd = P.init_state_dict(P.model_param_specs(D), seed=0)
    enc = _enc(D, sd).eval()
    g = torch.Generator().manual_seed(5)
    B = 256
    emb = torch.randn(B, 20, D.hidden, generator=g).cuda()
    feats = torch.randn(B, 64, D.feat_dim, generator=g).abs().cuda()
    pos = torch.rand(B, 64, 4, generator=g).cuda()
    with torch.no_grad():
        (v, _), (l, _), _ = enc(emb, None, feats, pos)
        (v3, _), (l3, _), _ = enc(emb[:3].contiguous(), None, feats[:3].contiguous(), pos[:3].contiguous())
    assert torch.isfinite(l[-1]).all() and torch.isfinite(v[-1]).all()
    assert rel_err(l[-1][:3].cpu(), l3[-1].cpu()) < 1e-6
    assert rel_err(v[-1][:3].cpu(), v3[-1].cpu()) < 1e-6


@pytest.mark.parametrize("B", [1, 3])
def test_generator_odd_batches_match_oracle(B):
    from xlxmert_b200.generator import B200Generator
    sd = P.init_generator_state_dict(seed=3)
    G = B200Generator()
    G.load_state_dict(sd, strict=True)
    G = G.cuda().eval()
    g = torch.Generator().manual_seed(B)
    emb = 0.05 * torch.randn(B, 2048, 8, 8, generator=g).abs()
    ref = GO.generator(sd, emb, return_intermediates=True)
    img, pre, hs = G(emb.cuda(), train=False, return_intermediates=True)
    ref_img = ref[0] if isinstance(ref, tuple) else ref
    assert float((img.cpu() - ref_img).abs().max()) < 1e-3
    assert img.shape == (B, 3, 256, 256)


def test_heads_large_ragged_rows():
    """Cluster head on a row count that is not a multiple of any tile size, with every label ignored except one."""
    from xlxmert_b200.heads import B200LxmertVisualObjHead
    d = TINY_DIMS
    sdh = P.init_state_dict(P.objhead_param_specs(d), seed=4, randomize_ln_bias=True)
    head = B200LxmertVisualObjHead(d, d.num_clusters)
    head.load_state_dict(sdh, strict=True)
    head = head.cuda()
    g = torch.Generator().manual_seed(1)
    h = torch.randn(3, 37, d.hidden, generator=g)
    labels = torch.full((3, 37), -100, dtype=torch.int64)
    labels[1, 5] = 17
    labels[2, 36] = d.num_clusters - 1
    sdo = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    ho = h.clone().requires_grad_(True)
    loss_o = O.obj_loss(sdo, ho, labels)
    loss_o.backward()
    hg = h.cuda().requires_grad_(True)
    loss = head.loss(hg, labels.cuda())
    loss.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    assert rel_err(hg.grad.cpu(), ho.grad) < 1e-3
    assert rel_err(head.out_cluster.bias.grad.cpu(), sdo["out_cluster.bias"].grad) < 1e-3
    # all labels ignored → NaN loss like torch's CrossEntropyLoss (0/0), never a crash
    nan_loss = head.loss(h.cuda(), torch.full((3, 37), -100, dtype=torch.int64).cuda())
    assert torch.isnan(nan_loss)


@pytest.mark.parametrize("M,frac", [(1, 1.0), (37, 0.0), (1000, 0.5), (16384, 0.15), (5000, 1.0)])
def test_labelled_rows_compaction(M, frac):
    """xlx_labelled_rows / gather / scatter against torch.nonzero: ascending order, exact count, adjoint pair."""
    import ctypes as C
    from xlxmert_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M)
    labels = torch.randint(0, 100, (M,), generator=g)
    labels[torch.rand(M, generator=g) >= frac] = -100
    want = (labels != -100).nonzero().squeeze(1)
    lab = labels.cuda()
    rows = torch.full((M,), -1, dtype=torch.int64, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.xlx_labelled_rows(lab.data_ptr(), M, -100, rows.data_ptr(), count.data_ptr(), st) == 0
    n = int(count.item())
    assert n == want.numel()
    assert torch.equal(rows[:n].cpu(), want)
    if n == 0:
        return
    x = torch.randn(M, 64, generator=g).cuda()
    picked = torch.empty(n, 64, device="cuda")
    plab = torch.empty(n, dtype=torch.int64, device="cuda")
    assert lib.xlx_gather_rows(x.data_ptr(), lab.data_ptr(), rows.data_ptr(), n, 64, picked.data_ptr(),
                               plab.data_ptr(), st) == 0
    assert torch.equal(picked.cpu(), x.cpu()[want]) and torch.equal(plab.cpu(), labels[want])
    back = torch.full((M, 64), 7.0, device="cuda")
    assert lib.xlx_scatter_rows(picked.data_ptr(), rows.data_ptr(), n, M, 64, back.data_ptr(), st) == 0
    ref = torch.zeros(M, 64)
    ref[want] = x.cpu()[want]
    assert torch.equal(back.cpu(), ref)


def test_head_losses_identical_with_and_without_row_compaction():
    """Dropping the −100 rows before the heads must not change the loss or any gradient beyond summation order."""
    from xlxmert_b200.heads import B200LxmertVisualObjHead
    d = TINY_DIMS
    sdh = P.init_state_dict(P.objhead_param_specs(d), seed=9, randomize_ln_bias=True)
    g = torch.Generator().manual_seed(3)
    h = torch.randn(5, 36, d.hidden, generator=g)
    labels = torch.randint(0, d.num_clusters, (5, 36), generator=g)
    labels[torch.rand(5, 36, generator=g) < 0.5] = -100
    res = []
    for compact in (False, True):
        head = B200LxmertVisualObjHead(d, d.num_clusters)
        head.load_state_dict(sdh, strict=True)
        head = head.cuda()
        head.compact_rows = compact
        hg = h.cuda().requires_grad_(True)
        loss = head.loss(hg, labels.cuda())
        loss.backward()
        res.append((float(loss), hg.grad.cpu(), {k: p.grad.cpu() for k, p in head.named_parameters()
                                                 if p.grad is not None}))
    (l0, g0, p0), (l1, g1, p1) = res
    assert abs(l0 - l1) < 1e-6 * abs(l0)
    assert g1.shape == g0.shape and rel_err(g1, g0) < 1e-5
    assert float(g1[labels == -100].abs().max()) == 0.0
    assert p0.keys() == p1.keys()
    for k in p0:
        assert rel_err(p1[k], p0[k]) < 1e-5, k
