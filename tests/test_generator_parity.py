"""GPU parity of the generator forward against the golden produced by the reference's own ``Generator``
(oracle/make_golden.py::golden_generator; eval mode, noise off, converged spectral-norm buffers).  Compared:
every residual block's output, the pre-tanh accumulator and the image (not the image alone —
tanh saturation hides errors).  Tolerance: 1e-3 tensor-normalised (north_star), observed far below."""
import pytest
import torch

from xlxmert_b200 import params as P
from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D

from util import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _setup():
    from xlxmert_b200.generator import B200Generator
    g = load_golden("generator_b2")
    B, wseed, bseed = (int(x) for x in g["meta"])
    G = B200Generator(base_dim=32, emb_dim=2048, norm_type="spade_in", target_size=256, init_H=8, init_W=8, SN=True,
                      codebook_dim=256)
    sd = P.init_generator_state_dict(seed=wseed)
    missing, unexpected = G.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    G = G.cuda().eval()
    batch = synth.make_batch(D, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"]).cuda()      # [B, 64, 2048]
    return g, G, code, B


def test_generator_matches_reference_golden():
    g, G, code, B = _setup()
    emb = code.permute(0, 2, 1).view(B, 2048, 8, 8)        # exactly what the sampler passes (imggen_model.py:254)
    img, pre, hs = G(emb, train=False, return_intermediates=True)
    assert img.shape == (B, 3, 256, 256)
    for i, h in enumerate(hs):
        s = max(1, h.shape[-1] // 16)
        e = rel_err(h[:, :, ::s, ::s].cpu(), g[f"h{i}_sub"])
        assert e < 1e-3, (i, e)
        assert abs(float(h.abs().mean()) - float(g[f"h{i}_absmean"])) < 1e-3 * float(g[f"h{i}_absmean"])
    assert rel_err(pre[:, :, ::4, ::4].cpu(), g["pre_tanh_sub"]) < 1e-3
    assert float((img[:, :, ::4, ::4].cpu() - torch.from_numpy(g["img_sub"])).abs().max()) < 1e-3
    assert abs(float(img.mean()) - float(g["img_mean"])) < 1e-4
    # the golden is not a vacuous target: most pixels are away from tanh saturation
    assert float(g["saturated_frac"]) < 0.5


def test_generator_input_layouts_and_noise_flag_agree():
    g, G, code, B = _setup()
    a = G(code.permute(0, 2, 1).view(B, 2048, 8, 8), train=False)
    b = G(code.view(B, 8, 8, 2048), train=False)            # layers.py:231-233 accepts cell-major input too
    c = G(code.permute(0, 2, 1).view(B, 2048, 8, 8).contiguous(), train=True)   # noise weights are 0 ⇒ identical
    assert torch.equal(a, b) and torch.equal(a, c)
    with torch.no_grad():
        for rb in G.resblocks:                               # a trained G has non-zero noise weights
            rb.noise1.weight.fill_(0.05)    # (in place on the Parameter: `.data.fill_` would not bump its version counter)
    d = G(code.view(B, 8, 8, 2048), train=True)
    assert not torch.equal(a, d) and torch.isfinite(d).all() and float(d.abs().max()) <= 1.0
    assert torch.equal(a, G(code.view(B, 8, 8, 2048), train=False))


def test_generator_rejects_other_architectures_and_cpu():
    from xlxmert_b200.generator import B200Generator
    with pytest.raises(NotImplementedError):
        B200Generator(base_dim=64)
    G = B200Generator()
    with pytest.raises(RuntimeError):
        G(torch.zeros(1, 2048, 8, 8))


def test_single_pass_bf16_mode_tracks_the_default_mode():
    """``passes=1`` (plain bf16 operands, hi parts only) runs the same kernels with one operand part per stage — the
    row-halo convolution and its rank-3 weight box included.  Not a parity mode: it must stay close, not equal."""
    from xlxmert_b200.generator import B200Generator
    g, G, code, B = _setup()
    G1 = B200Generator(passes=1)
    G1.load_state_dict(G.state_dict(), strict=True)
    G1 = G1.cuda().eval()
    ref, pre_ref, _ = G(code.view(B, 8, 8, 2048), train=False, return_intermediates=True)
    out, pre, _ = G1(code.view(B, 8, 8, 2048), train=False, return_intermediates=True)
    assert torch.isfinite(out).all()
    assert rel_err(pre.cpu(), pre_ref.cpu()) < 5e-2
    assert float((out - ref).abs().mean()) < 1e-2


def _generator(sd=None, wseed=0):
    from xlxmert_b200.generator import B200Generator
    G = B200Generator()
    G.load_state_dict(sd if sd is not None else P.init_generator_state_dict(seed=wseed), strict=True)
    return G.cuda().eval()


def test_generator_batch_16_matches_reference_golden():
    """B = 16: at the 8×8 and 16×16 stages a 128-pixel GEMM tile spans two images / half an image, cluster pairing and
    wave counts differ from B = 2 — every image of the batch is compared with the reference's own run
    (golden ``generator_b16``, oracle/make_golden.py::golden_generator_batch)."""
    g = load_golden("generator_b16")
    B, wseed, bseed = (int(x) for x in g["meta"])
    G = _generator(wseed=wseed)
    batch = synth.make_batch(D, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"]).cuda()
    img, pre, hs = G(code.permute(0, 2, 1).view(B, 2048, 8, 8), train=False, return_intermediates=True)
    keep = [int(k) for k in g["keep"]]
    for i, h in enumerate(hs):
        s = max(1, h.shape[-1] // 16)
        e = rel_err(h[keep][:, :, ::s, ::s].cpu(), g[f"h{i}_sub"])
        assert e < 1e-3, (i, e)
    for b in range(B):      # per image, so that one bad image cannot hide behind the batch maximum
        assert rel_err(pre[b, :, ::8, ::8].cpu(), g["pre_tanh_sub"][b]) < 1e-3, b
        assert float((img[b, :, ::8, ::8].cpu() - torch.from_numpy(g["img_sub"][b])).abs().max()) < 1e-3, b
    assert float((img.mean(dim=(1, 2, 3)).cpu() - torch.from_numpy(g["img_mean_per_image"])).abs().max()) < 1e-4


def test_generator_fixed_noise_matches_reference_golden():
    """``forward(train=True)`` of a G with non-zero noise weights (layers.py:56-62): the reference draws
    ``new_empty(B,1,R,R).normal_()`` from the global CPU generator, noise1 then noise2 of each block; re-drawing the same
    maps from the recorded seed and injecting them must reproduce the reference's run."""
    g = load_golden("generator_noise_b2")
    B, wseed, bseed, noise_seed = (int(x) for x in g["meta"])
    sd = P.init_generator_state_dict(seed=wseed)
    for k in sd:
        if k.endswith("noise1.weight") or k.endswith("noise2.weight"):
            sd[k] = torch.full_like(sd[k], float(g["noise_weight"]))
    G = _generator(sd)
    batch = synth.make_batch(D, B, 20, 64, seed=bseed)
    code = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"]).cuda()
    torch.manual_seed(noise_seed)
    noise = []
    for i in range(5):
        noise += [torch.empty(B, 1, 8 << i, 8 << i).normal_(), torch.empty(B, 1, 16 << i, 16 << i).normal_()]
    emb = code.permute(0, 2, 1).view(B, 2048, 8, 8)
    img, pre, hs = G(emb, train=True, return_intermediates=True, noise=noise)
    for i, h in enumerate(hs):
        s = max(1, h.shape[-1] // 16)
        assert rel_err(h[:, :, ::s, ::s].cpu(), g[f"h{i}_sub"]) < 1e-3, i
    assert rel_err(pre[:, :, ::4, ::4].cpu(), g["pre_tanh_sub"]) < 1e-3
    assert float((img[:, :, ::4, ::4].cpu() - torch.from_numpy(g["img_sub"])).abs().max()) < 1e-3
    # and the noise really took part: the noise-free forward differs
    assert not torch.equal(img, G(emb, train=False))


def test_generator_batch_128_subsample_against_oracle():
    """BASELINE.json configs[3] size (B = 128): a few images of the full-size batch (first / middle / last — different
    tiles, waves and cluster pairs) against the CPU oracle on the same codes.  InstanceNorm makes every image independent
    of its batch (layers.py:33-47), so the oracle runs on the sub-sample only."""
    from oracle import generator_oracle as GO
    B = 128
    sd = P.init_generator_state_dict(seed=0)
    G = _generator(sd)
    batch = synth.make_batch(D, B, 20, 64, seed=21)
    code = synth.visual_feats_from(synth.centroid_table(D), batch["cluster_ids"])
    img, pre, _ = G(code.cuda().view(B, 8, 8, 2048), train=False, return_intermediates=True)
    pick = [0, 63, 127]
    with torch.no_grad():
        ref_img, inter = GO.generator(sd, code[pick].permute(0, 2, 1).reshape(len(pick), 2048, 8, 8),
                                      return_intermediates=True)
    ref_pre = inter["pre_tanh"]
    for j, b in enumerate(pick):
        assert rel_err(pre[b].cpu(), ref_pre[j]) < 1e-3, b
        assert float((img[b].cpu() - ref_img[j]).abs().max()) < 1e-3, b
    # batch independence at full size: the same three images alone give bit-identical results
    alone = G(code[pick].cuda().view(len(pick), 8, 8, 2048), train=False)
    assert torch.equal(alone, img[pick])
