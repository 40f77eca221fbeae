"""Nearest-centroid assignment with the call shape of ``faiss.IndexFlatL2`` as the reference uses it
(``x-lxmert/feature_extraction/run_kmeans.py:124-143``)::

    index = B200IndexFlatL2(d); index.add(centroids); D, I = index.search(x, 1)

``x`` may be a numpy array or CPU tensor of any length (streamed through two pinned staging buffers on two CUDA
streams so the host→device copy of one chunk overlaps the GEMM of the previous one) or a CUDA tensor.  Returns what
faiss returns: ``D`` float32 ``[N, 1]`` squared L2 distances, ``I`` int64 ``[N, 1]`` centroid ids — numpy for
numpy/CPU input, CUDA tensors for CUDA input.  Only ``k == 1`` exists; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class B200IndexFlatL2:
    def __init__(self, d: int, device="cuda", passes: int = 3, chunk_rows: int = 32768):
        self.d, self.passes, self.chunk_rows = int(d), passes, int(chunk_rows)
        self.device = torch.device(device)
        self._centroids = None
        self._prep = None
        self._ws = {}

    @property
    def ntotal(self) -> int:
        return 0 if self._centroids is None else self._centroids.shape[0]

    def add(self, centroids):
        c = torch.as_tensor(np.ascontiguousarray(centroids) if isinstance(centroids, np.ndarray) else centroids)
        if c.dim() != 2 or c.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] vectors, got {tuple(c.shape)}")
        if self.device.type != "cuda":
            raise RuntimeError("B200IndexFlatL2 runs on CUDA (sm_100a) only; there is no CPU fallback")
        c = c.to(self.device, torch.float32).contiguous()
        self._centroids = c if self._centroids is None else torch.cat([self._centroids, c])
        lib = _lib.load()
        K = self.ntotal
        n = lib.xlx_kmeans_prep_bytes(self.d, K)
        if n == 0:
            raise _lib.XlxError("xlx_kmeans_prep_bytes", -20)
        self._prep = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.xlx_kmeans_prepare(self.d, K, self._centroids.data_ptr(), self._prep.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
        _lib.check("xlx_kmeans_prepare", rc)

    def _workspace(self, slot, N):
        lib = _lib.load()
        need = lib.xlx_kmeans_workspace_bytes(self.d, self.ntotal, N)
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < need:
            ws = self._ws[slot] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def _assign(self, x, ids, dist, slot, stream):
        N = x.shape[0]
        ws = self._workspace(slot, N)
        rc = _lib.load().xlx_kmeans_assign(self.d, self.ntotal, self._prep.data_ptr(), N, x.data_ptr(), ids.data_ptr(),
                                           dist.data_ptr(), ws.data_ptr(), ws.numel(), self.passes, stream.cuda_stream)
        _lib.check("xlx_kmeans_assign", rc)

    @torch.no_grad()
    def search(self, x, k: int = 1):
        if k != 1:
            raise NotImplementedError("only the nearest centroid (k = 1) is implemented (run_kmeans.py:143)")
        if self._prep is None:
            raise RuntimeError("search() before add()")
        on_gpu = isinstance(x, torch.Tensor) and x.is_cuda
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] queries, got {tuple(x.shape)}")
        N = x.shape[0]
        with torch.cuda.device(self.device):
            if on_gpu:
                x = x.float().contiguous()
                ids = torch.empty(N, dtype=torch.int64, device=self.device)
                dist = torch.empty(N, dtype=torch.float32, device=self.device)
                for r0 in range(0, N, self.chunk_rows):
                    r1 = min(N, r0 + self.chunk_rows)
                    self._assign(x[r0:r1], ids[r0:r1], dist[r0:r1], 0, torch.cuda.current_stream())
                return dist.view(N, 1), ids.view(N, 1)
            return self._search_host(x.float().contiguous(), N)

    def _search_host(self, x, N):
        R = min(self.chunk_rows, max(N, 1))
        ids_out = torch.empty(N, dtype=torch.int64).pin_memory()
        dist_out = torch.empty(N, dtype=torch.float32).pin_memory()
        stage = [torch.empty(R, self.d, dtype=torch.float32).pin_memory() for _ in range(2)]
        dev_x = [torch.empty(R, self.d, dtype=torch.float32, device=self.device) for _ in range(2)]
        dev_i = [torch.empty(R, dtype=torch.int64, device=self.device) for _ in range(2)]
        dev_d = [torch.empty(R, dtype=torch.float32, device=self.device) for _ in range(2)]
        streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        for st in streams:       # add() queued the centroid upload + xlx_kmeans_prepare on the current stream
            st.wait_stream(torch.cuda.current_stream(self.device))
        busy = [None, None]
        for i, r0 in enumerate(range(0, N, R)):
            s = i & 1
            r1 = min(N, r0 + R)
            n = r1 - r0
            if busy[s] is not None:
                busy[s].synchronize()              # staging buffer s is free again
            stage[s][:n].copy_(x[r0:r1])
            with torch.cuda.stream(streams[s]):
                dev_x[s][:n].copy_(stage[s][:n], non_blocking=True)
                self._assign(dev_x[s][:n], dev_i[s][:n], dev_d[s][:n], s, streams[s])
                ids_out[r0:r1].copy_(dev_i[s][:n], non_blocking=True)
                dist_out[r0:r1].copy_(dev_d[s][:n], non_blocking=True)
                busy[s] = torch.cuda.Event()
                busy[s].record(streams[s])
        for e in busy:
            if e is not None:
                e.synchronize()
        return dist_out.numpy().reshape(N, 1).copy(), ids_out.numpy().reshape(N, 1).copy()
