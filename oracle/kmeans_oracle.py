"""TEST INFRASTRUCTURE ONLY — CPU restatement of the nearest-centroid search the reference performs with
``faiss.IndexFlatL2(d).search(x, 1)`` (``x-lxmert/feature_extraction/run_kmeans.py:124-143``).

PARITY UNPINNED: faiss is a third-party dependency that is not vendored in the reference and not installed here
(the reference's README pins nothing; faiss-cpu 1.6.x was current at its release), and the reference ships no golden
assignments.  This restates faiss's published algorithm for the flat L2 index (``knn_L2sqr`` BLAS path: distances as
‖x‖² + ‖c‖² − 2·x·cᵀ from an SGEMM, negative round-off clamped to 0, the smallest distance kept, lowest index on
ties) in fp32, plus an fp64 brute-force version used to tell real disagreements from near-ties.
tests/test_kmeans.py cross-checks both functions against scikit-learn's pairwise_distances_argmin_min — an independent
implementation of the same definition, not the reference's dependency, so the header stays "unpinned".
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def search_l2_fp32(x: np.ndarray, centroids: np.ndarray):
    """→ (D [N,1] float32, I [N,1] int64) the way faiss's flat index computes them."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    xn = (x * x).sum(1, dtype=np.float32)
    cn = (c * c).sum(1, dtype=np.float32)
    d = xn[:, None] + cn[None, :] - np.float32(2.0) * (x @ c.T)
    np.maximum(d, 0, out=d)
    i = d.argmin(1)
    return d[np.arange(len(x)), i].reshape(-1, 1), i.astype(np.int64).reshape(-1, 1)


def search_l2_fp64(x: np.ndarray, centroids: np.ndarray):
    """Exact (fp64) distances: → (best distance, best id, margin to the runner-up) per row."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(centroids, dtype=np.float64)
    d = (x * x).sum(1)[:, None] + (c * c).sum(1)[None, :] - 2.0 * (x @ c.T)
    i = d.argmin(1)
    rows = np.arange(len(x))
    best = d[rows, i]
    if c.shape[0] > 1:
        d2 = d.copy()
        d2[rows, i] = np.inf
        margin = d2.min(1) - best
    else:
        margin = np.full(len(x), np.inf)
    return best, i.astype(np.int64), margin
