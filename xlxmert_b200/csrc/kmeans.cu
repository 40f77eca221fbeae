// Nearest-centroid assignment — §8(f): the reference's feature_extraction/run_kmeans.py:124-143 builds a
// faiss.IndexFlatL2 over the [10000, 2048] centroids and calls index.search(x, 1) for every grid cell of every image.
// Here: argmin_c ‖x − c‖² = argmax_c (2·x·c − ‖c‖²).  The centroid table is prepared once as split-bf16(2·C) with
// −‖c‖² as the epilogue bias, so a chunk of rows costs one fused row pass (‖x‖² + bf16 split), one tcgen05 GEMM
// (same shape as the cluster head's out_cluster layer) and one row arg-max.
#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

namespace xlx {
namespace {

inline int pad8(int n) { return (n + 7) & ~7; }

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, size_t idx, float4 v) {
  bf16 h[4], l[4];
  const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __float2bfloat16_rn(f[i]);
    l[i] = __float2bfloat16_rn(f[i] - __bfloat162float(h[i]));
  }
  *reinterpret_cast<uint2*>(hi + idx) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lo + idx) = *reinterpret_cast<const uint2*>(l);
}

// one CTA per row: out = split(scale · row), norm[row] = sign · ‖row‖²
__global__ void __launch_bounds__(256) row_split_norm_kernel(const float* __restrict__ x, int cols, float scale,
                                                             float sign, bf16* hi, bf16* lo, float* norm) {
  __shared__ float sh[8];
  const size_t r = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(x + r * cols);
  float acc = 0.f;
  for (int c = threadIdx.x; c < cols / 4; c += 256) {
    float4 v = __ldg(src + c);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    store_split(hi, lo, r * cols + c * 4, make_float4(scale * v.x, scale * v.y, scale * v.z, scale * v.w));
  }
  acc = block_sum_256(acc, sh);
  if (threadIdx.x == 0) norm[r] = sign * acc;
}

// one CTA per row: id = first index of the row maximum over C columns; dist = max(xnorm − best, 0)
__global__ void __launch_bounds__(256) row_argmax_kernel(const float* __restrict__ score, int ld, int C,
                                                         const float* __restrict__ xnorm, int64_t* __restrict__ ids,
                                                         float* __restrict__ dist) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const size_t r = blockIdx.x;
  const float* row = score + r * ld;
  float best = -INFINITY;
  int at = 0x7fffffff;
  for (int c = threadIdx.x; c < C; c += 256) {      // ascending c per thread: strict > keeps the first index
    float v = row[c];
    if (v > best) { best = v; at = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, at, o);
    if (ov > best || (ov == best && oi < at)) { best = ov; at = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = at; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < at)) { best = sv[w]; at = si[w]; }
    ids[r] = at;
    if (dist) dist[r] = fmaxf(xnorm[r] - best, 0.f);
  }
}

struct KmPrep {
  Split w;        // [Kp, dim] = split(2·centroids), rows K..Kp-1 unused (the GEMM treats them as zero)
  float* bias;    // [Kp] = −‖c‖²
  size_t bytes;
};
KmPrep km_prep(int dim, int K, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  KmPrep p;
  p.w = b.split(static_cast<size_t>(pad8(K)) * dim);
  p.bias = b.f32(pad8(K));
  p.bytes = b.total();
  return p;
}
struct KmWs {
  Split x;
  float* xnorm;
  float* score;
  size_t bytes;
};
KmWs km_ws(int dim, int K, int N, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  KmWs w;
  w.x = b.split(static_cast<size_t>(N) * dim);
  w.xnorm = b.f32(N);
  w.score = b.f32(static_cast<size_t>(N) * pad8(K));
  w.bytes = b.total();
  return w;
}
inline bool km_ok(int dim, int K) { return dim >= 32 && dim % 8 == 0 && K >= 1; }

}  // namespace
}  // namespace xlx

using namespace xlx;

extern "C" {

size_t xlx_kmeans_prep_bytes(int32_t dim, int32_t n_centroids) {
  return km_ok(dim, n_centroids) ? km_prep(dim, n_centroids, nullptr).bytes : 0;
}
int32_t xlx_kmeans_prepare(int32_t dim, int32_t n_centroids, const float* centroids, void* prep, void* stream) {
  if (!km_ok(dim, n_centroids)) return -20;
  if (!centroids || !prep) return -24;
  XLX_TRY(ensure_device(prep));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KmPrep p = km_prep(dim, n_centroids, prep);
  XLX_CUDA(cudaMemsetAsync(p.bias, 0, static_cast<size_t>(pad8(n_centroids)) * 4, st));
  row_split_norm_kernel<<<n_centroids, 256, 0, st>>>(centroids, dim, 2.0f, -1.0f, p.w.hi, p.w.lo, p.bias);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}
size_t xlx_kmeans_workspace_bytes(int32_t dim, int32_t n_centroids, int32_t N) {
  return (km_ok(dim, n_centroids) && N > 0) ? km_ws(dim, n_centroids, N, nullptr).bytes : 0;
}
int32_t xlx_kmeans_assign(int32_t dim, int32_t n_centroids, const void* prep, int32_t N, const float* x, int64_t* ids,
                          float* dist, void* workspace, size_t workspace_bytes, int32_t passes, void* stream) {
  if (!km_ok(dim, n_centroids)) return -20;
  if (N < 1) return -21;
  if (!prep || !x || !ids || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KmPrep p = km_prep(dim, n_centroids, const_cast<void*>(prep));
  KmWs w = km_ws(dim, n_centroids, N, workspace);
  if (w.bytes > workspace_bytes) return -23;
  const int Kp = pad8(n_centroids);
  row_split_norm_kernel<<<N, 256, 0, st>>>(x, dim, 1.0f, 1.0f, w.x.hi, w.x.lo, w.xnorm);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  GemmEpilogue e;
  e.bias = p.bias; e.out_f32 = w.score; e.ld_out = Kp;
  XLX_TRY(gemm_linear(passes, st, w.x, N, dim, p.w, Kp, e, n_centroids));
  row_argmax_kernel<<<N, 256, 0, st>>>(w.score, Kp, n_centroids, w.xnorm, ids, dist);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
