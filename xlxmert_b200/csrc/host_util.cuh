// Host-side helpers shared by the C-ABI translation units: error plumbing, a bump allocator over
// caller-owned workspaces, and the three GEMM shapes every nn.Linear needs (forward, dgrad, wgrad).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "gemm_sm100.cuh"
#include "kernels.cuh"

namespace xlx {

#define XLX_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__) return rc__;       \
  } while (0)

#define XLX_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t e__ = (expr);                                 \
    if (e__ != cudaSuccess) return static_cast<int>(e__);     \
  } while (0)

// Carves 256-byte aligned pieces out of a caller-owned buffer (base may be null to only measure).
struct Bump {
  char* base = nullptr;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base + off;
    off += bytes;
    return p;
  }
  float* f32(size_t n) { return static_cast<float*>(take(n * 4)); }
  Split split(size_t n) {
    Split s;
    s.hi = static_cast<bf16*>(take(n * 2));
    s.lo = static_cast<bf16*>(take(n * 2));
    return s;
  }
  size_t total() const { return off + 256; }
};
inline Split rows(Split s, size_t row0, size_t ld) {
  Split r;
  r.hi = s.hi + row0 * ld;
  r.lo = s.lo ? s.lo + row0 * ld : nullptr;
  return r;
}

// The caller (PyTorch) may have selected the device through a different copy of the CUDA runtime; make this
// library's runtime agree with the device that owns the caller's buffers.
inline int ensure_device(const void* ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (at.type != cudaMemoryTypeDevice) return 0;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != at.device) {
    cudaError_t e = cudaSetDevice(at.device);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
}

// Y[M,N] = X[M,K] · W[N,K]ᵀ (+ epilogue); ldx = leading dimension of X (0 → K)
// w_rows: rows of W that exist (0 → N); N may be padded beyond it (extra output columns see zero weights)
inline int gemm_linear(int passes, cudaStream_t st, Split x, int M, int K, Split w, int N, const GemmEpilogue& e,
                       int w_rows = 0) {
  GemmProblem p;
  p.M = M; p.N = N; p.K = K; p.passes = passes;
  p.a.hi = x.hi; p.a.lo = x.lo; p.a.ld = K; p.a.mn_major = 0;
  p.b.hi = w.hi; p.b.lo = w.lo; p.b.ld = K; p.b.mn_major = 0; p.b.rows = w_rows;
  p.epi = e;
  return gemm_launch(p, st);
}
// dX[M,K] = dY[M,N] · W[N,K];  w_rows: rows of W that exist (0 → N; dY columns beyond must be finite)
inline int gemm_dgrad(int passes, cudaStream_t st, Split dy, int M, int N, Split w, int K, const GemmEpilogue& e,
                      int w_rows = 0) {
  GemmProblem p;
  p.M = M; p.N = K; p.K = N; p.passes = passes;
  p.a.hi = dy.hi; p.a.lo = dy.lo; p.a.ld = N; p.a.mn_major = 0;
  p.b.hi = w.hi; p.b.lo = w.lo; p.b.ld = K; p.b.mn_major = 1;   // W stored [N, K]: GEMM-N (= K) contiguous
  p.b.kext = w_rows;
  p.epi = e;
  return gemm_launch(p, st);
}
// dW[N,K] (+)= dY[M,N]ᵀ · X[M,K];  ld_dy: leading dimension of dY (0 → N)
inline int gemm_wgrad(int passes, cudaStream_t st, Split dy, int M, int N, Split x, int K, float* dw,
                      bool accumulate = false, int ld_dy = 0, float* splitk_ws = nullptr) {
  GemmProblem p;
  p.M = N; p.N = K; p.K = M; p.passes = passes;
  p.a.hi = dy.hi; p.a.lo = dy.lo; p.a.ld = ld_dy ? ld_dy : N; p.a.mn_major = 1;
  p.b.hi = x.hi; p.b.lo = x.lo; p.b.ld = K; p.b.mn_major = 1;
  p.epi.out_f32 = dw; p.epi.ld_out = K;
  if (accumulate) p.epi.flags |= EPI_ACCUM;
  p.splitk_ws = splitk_ws; p.splitk_ws_floats = splitk_ws ? gemm_splitk_ws_floats() : 0;
  return gemm_launch(p, st);
}

}  // namespace xlx
