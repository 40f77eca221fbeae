"""Nearest-centroid assignment (run_kmeans.py:124-143) against the fp32 / fp64 oracle restatements of
faiss.IndexFlatL2.search(x, 1).  Index parity is exact wherever the fp64 margin between the two closest centroids
exceeds the fp32 round-off of the distance itself; distances within 1e-4 relative (of ‖x‖²)."""
import numpy as np
import pytest
import torch

from oracle import kmeans_oracle as KO


def _data(N, K, d, seed):
    rng = np.random.default_rng(seed)
    cent = np.abs(rng.standard_normal((K, d))).astype(np.float32)          # ResNet features are post-ReLU
    pick = rng.integers(0, K, N)
    x = (cent[pick] + 0.35 * rng.standard_normal((N, d))).astype(np.float32)
    return x, cent


def test_oracle_fp32_agrees_with_fp64():
    x, c = _data(300, 97, 64, 0)
    D, I = KO.search_l2_fp32(x, c)
    best, i64, margin = KO.search_l2_fp64(x, c)
    safe = margin > 1e-3
    assert safe.mean() > 0.9
    assert np.array_equal(I[safe, 0], i64[safe])
    assert np.allclose(D[:, 0], best, rtol=1e-4, atol=1e-3)
    # ties go to the lowest index; an exact hit has distance 0 (clamped, never negative)
    c2 = np.concatenate([c, c[:5]])
    D2, I2 = KO.search_l2_fp32(c[:5], c2)
    assert np.array_equal(I2[:, 0], np.arange(5)) and (D2 >= 0).all() and (D2 < 1e-3).all()


def test_oracle_agrees_with_scikit_learn():
    """faiss is absent, so the restatement is cross-checked against an independent implementation of the same
    definition (nearest centroid under squared L2): scikit-learn's pairwise_distances_argmin_min."""
    sk = pytest.importorskip("sklearn.metrics")
    x, c = _data(500, 211, 96, 7)
    idx, dist = sk.pairwise_distances_argmin_min(x.astype(np.float64), c.astype(np.float64), metric="sqeuclidean")
    best, i64, margin = KO.search_l2_fp64(x, c)
    assert np.array_equal(i64, idx)
    assert np.allclose(best, dist, rtol=1e-9, atol=1e-9)
    D, I = KO.search_l2_fp32(x, c)
    safe = margin > 1e-3
    assert np.array_equal(I[safe, 0], idx[safe]) and np.allclose(D[:, 0], dist, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,d,chunk", [(1, 1, 32, 64), (257, 97, 64, 100), (1000, 1001, 256, 4096),
                                          (4096, 10000, 2048, 1536)])
def test_assignment_matches_oracle(N, K, d, chunk):
    from xlxmert_b200.kmeans import B200IndexFlatL2
    x, c = _data(N, K, d, N + K)
    index = B200IndexFlatL2(d, chunk_rows=chunk)
    index.add(c)
    assert index.ntotal == K
    D, I = index.search(x, 1)                    # numpy in → numpy out, streamed in chunks
    assert D.shape == (N, 1) and I.shape == (N, 1) and D.dtype == np.float32 and I.dtype == np.int64
    best, i64, margin = KO.search_l2_fp64(x, c)
    scale = (x.astype(np.float64) ** 2).sum(1)
    safe = margin > 1e-5 * scale
    assert np.array_equal(I[safe, 0], i64[safe])
    # a near-tie may pick the runner-up, but never a centroid that is measurably farther
    chosen = ((x.astype(np.float64) - c[I[:, 0]].astype(np.float64)) ** 2).sum(1)
    assert (chosen - best <= 1e-5 * scale + 1e-9).all()
    assert np.abs(D[:, 0] - best).max() <= 1e-4 * scale.max()
    Dt, It = index.search(torch.from_numpy(x).cuda(), 1)       # CUDA in → CUDA out
    assert It.is_cuda and torch.equal(It.cpu().view(-1), torch.from_numpy(I[:, 0]))
    assert torch.equal(Dt.cpu().view(-1), torch.from_numpy(D[:, 0]))


@pytest.mark.gpu
def test_centroids_assign_to_themselves_at_full_size():
    """Size-independent property at the reference's table size: search(C) = (≈0, arange)."""
    from xlxmert_b200.kmeans import B200IndexFlatL2
    _, c = _data(1, 10000, 2048, 5)
    index = B200IndexFlatL2(2048)
    index.add(c)
    D, I = index.search(torch.from_numpy(c).cuda(), 1)
    assert torch.equal(I.view(-1).cpu(), torch.arange(10000))
    norm = float((torch.from_numpy(c) ** 2).sum(1).max())
    assert float(D.min()) >= 0.0 and float(D.max()) < 1e-4 * norm
    with pytest.raises(NotImplementedError):
        index.search(torch.from_numpy(c[:4]).cuda(), 2)
