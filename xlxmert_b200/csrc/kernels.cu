// Non-GEMM kernels of the X-LXMERT hot path — see kernels.cuh.
#include "kernels.cuh"

#include <atomic>
#include <cstdlib>

#include "xlx_ptx.cuh"

namespace xlx {

namespace {

std::atomic<long long> g_aux_launches{0};
constexpr int kMaxBlocks = 296;  // 2 × 148 SMs: enough CTAs in flight for the HBM-bound reductions

inline int launch_rc() {
  g_aux_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// store 4 floats as split bf16 (8 bytes to hi, 8 bytes to lo)
__device__ __forceinline__ void store_split4(bf16* hi, bf16* lo, size_t idx, float4 v) {
  uint2 h, l;
  split_bf16x2(v.x, v.y, h.x, l.x);
  split_bf16x2(v.z, v.w, h.y, l.y);
  *reinterpret_cast<uint2*>(hi + idx) = h;
  if (lo) *reinterpret_cast<uint2*>(lo + idx) = l;
}
__device__ __forceinline__ float4 load_split4(const bf16* hi, const bf16* lo, size_t idx) {
  uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + idx));
  float4 v = make_float4(bf16lo_to_f32(h.x), bf16hi_to_f32(h.x), bf16lo_to_f32(h.y), bf16hi_to_f32(h.y));
  if (lo) {
    uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + idx));
    v.x += bf16lo_to_f32(l.x); v.y += bf16hi_to_f32(l.x); v.z += bf16lo_to_f32(l.y); v.w += bf16hi_to_f32(l.y);
  }
  return v;
}

// ---- elementwise -----------------------------------------------------------------------------
__global__ void split_kernel(const float4* __restrict__ x, bf16* hi, bf16* lo, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store_split4(hi, lo, i * 4, __ldg(x + i));
}
__global__ void unsplit_kernel(const bf16* hi, const bf16* lo, float4* out, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = load_split4(hi, lo, i * 4);
}
__global__ void fill_kernel(float* dst, float v, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = v;
}
__global__ void concat3_kernel(const float* a, const float* b, const float* c, float* dst, int n) {
  pdl_trigger();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * n) dst[i] = i < n ? a[i] : (i < 2 * n ? b[i - n] : c[i - 2 * n]);
}
// one CTA per output row
__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids,
                                   const uint8_t* __restrict__ mask, const float* __restrict__ fill, int cols,
                                   float* out_f32, bf16* hi, bf16* lo) {
  const int r = blockIdx.x;
  const bool m = mask && mask[r];
  const float4* src = reinterpret_cast<const float4*>(m ? fill : table + static_cast<size_t>(ids[r]) * cols);
  for (int c = threadIdx.x; c < cols / 4; c += blockDim.x) {
    float4 v = __ldg(src + c);
    size_t idx = static_cast<size_t>(r) * cols + c * 4;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + idx) = v;
    if (hi) store_split4(hi, lo, idx, v);
  }
}

// one CTA per row
__global__ void split_rows_kernel(const float* __restrict__ x, size_t ld_in, int cols, bf16* hi, bf16* lo) {
  const int r = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_in);
  for (int c = threadIdx.x; c < cols / 4; c += blockDim.x)
    store_split4(hi, lo, static_cast<size_t>(r) * cols + c * 4, __ldg(src + c));
}
__global__ void tanh_bwd_kernel(const float4* __restrict__ d, const float4* __restrict__ y, bf16* hi, bf16* lo,
                                size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(d + i), t = __ldg(y + i);
    store_split4(hi, lo, i * 4, make_float4(a.x * (1.f - t.x * t.x), a.y * (1.f - t.y * t.y),
                                            a.z * (1.f - t.z * t.z), a.w * (1.f - t.w * t.w)));
  }
}
// one CTA per token row
__global__ void embed_sum_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                 const float* __restrict__ word, const float* __restrict__ pos,
                                 const float* __restrict__ type, int L, int H, float* y) {
  const int r = blockIdx.x;
  // ids == nullptr: `word` is an inputs_embeds matrix [rows, H] instead of the table (HF:199-203)
  const float4* w = reinterpret_cast<const float4*>(word + static_cast<size_t>(ids ? ids[r] : r) * H);
  const float4* p = reinterpret_cast<const float4*>(pos + static_cast<size_t>(r % L) * H);
  const float4* t = reinterpret_cast<const float4*>(type + static_cast<size_t>(tt ? tt[r] : 0) * H);
  for (int c = threadIdx.x; c < H / 4; c += blockDim.x) {
    const float4 a = __ldg(w + c), b = __ldg(p + c), e = __ldg(t + c);
    // same association as torch: (word + position) + token_type
    *reinterpret_cast<float4*>(y + static_cast<size_t>(r) * H + c * 4) =
        make_float4((a.x + b.x) + e.x, (a.y + b.y) + e.y, (a.z + b.z) + e.z, (a.w + b.w) + e.w);
  }
}
__global__ void embed_scatter_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                     const float* __restrict__ dy, int L, int H, float* dword, float* dpos,
                                     float* dtype) {
  const int r = blockIdx.x;
  const int64_t id = ids ? ids[r] : 0, ty = tt ? tt[r] : 0;     // no ids (inputs_embeds): the word table gets no gradient
  const int l = r % L;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float g = __ldg(dy + static_cast<size_t>(r) * H + c);
    if (id != 0) atomicAdd(dword + static_cast<size_t>(id) * H + c, g);
    if (l != 0) atomicAdd(dpos + static_cast<size_t>(l) * H + c, g);
    if (ty != 0) atomicAdd(dtype + static_cast<size_t>(ty) * H + c, g);
  }
}

__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__global__ void gelu_bwd_kernel(const float4* __restrict__ dg, const float4* __restrict__ u, bf16* hi, bf16* lo,
                                size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(dg + i), x = __ldg(u + i);
    store_split4(hi, lo, i * 4, make_float4(a.x * gelu_grad_f(x.x), a.y * gelu_grad_f(x.y), a.z * gelu_grad_f(x.z),
                                            a.w * gelu_grad_f(x.w)));
  }
}

// ---- cross-entropy ------------------------------------------------------------------------------
constexpr int CE_T = 256;
// online (max, Σexp) over a row; every thread ends with the block-wide result
__device__ __forceinline__ void row_max_sumexp(const float* __restrict__ row, int C, float& mx_out, float& sum_out) {
  __shared__ float s_m[CE_T / 32], s_s[CE_T / 32];
  float m = -INFINITY, s = 0.f;
  for (int c = threadIdx.x * 4; c < C; c += CE_T * 4) {
    float v[4];
    if (c + 4 <= C) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(row + c));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (c + j < C) ? __ldg(row + c + j) : -INFINITY;
    }
    const float m2 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    if (m2 > m) { s *= expf(m - m2); m = m2; }
#pragma unroll
    for (int j = 0; j < 4; ++j) s += expf(v[j] - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * expf(m2 - mm));
    m = mm;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { s_m[w] = m; s_s[w] = s; }
  __syncthreads();
  float M = -INFINITY, S = 0.f;
#pragma unroll
  for (int i = 0; i < CE_T / 32; ++i) M = fmaxf(M, s_m[i]);
#pragma unroll
  for (int i = 0; i < CE_T / 32; ++i) S += (s_m[i] == -INFINITY) ? 0.f : s_s[i] * expf(s_m[i] - M);
  mx_out = M; sum_out = S;
}
__global__ void __launch_bounds__(CE_T)
ce_fwd_kernel(const float* __restrict__ logits, int ld, int C, const int64_t* __restrict__ labels, int64_t ignore,
              float* lse, float* rowloss) {
  const int m = blockIdx.x;
  const float* row = logits + static_cast<size_t>(m) * ld;
  float mx, sum;
  row_max_sumexp(row, C, mx, sum);
  if (threadIdx.x == 0) {
    const float l = mx + logf(sum);
    lse[m] = l;
    const int64_t y = labels[m];
    rowloss[m] = (y == ignore) ? 0.f : l - __ldg(row + y);
  }
}
// ---- feature regression loss (lxrt/modeling.py:270-284) -------------------------------------------------------------
// rowloss[r] = w[r] · Σ_f SmoothL1(feat[r,f] − target[r,f]),  β = 1: 0.5·d² for |d| < 1, |d| − 0.5 otherwise
// (torch.nn.SmoothL1Loss(reduction='none'), modeling.py:105).  One CTA per row, fixed reduction order.
__global__ void __launch_bounds__(256)
smooth_l1_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ target, const float* __restrict__ w, int F,
                     float* rowloss) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float wr = w[r];
  float acc = 0.f;
  if (wr != 0.f) {
    const float4* a = reinterpret_cast<const float4*>(feat + static_cast<size_t>(r) * F);
    const float4* b = reinterpret_cast<const float4*>(target + static_cast<size_t>(r) * F);
    for (int c = threadIdx.x; c < F / 4; c += blockDim.x) {
      const float4 x = __ldg(a + c), y = __ldg(b + c);
      const float d[4] = {fabsf(x.x - y.x), fabsf(x.y - y.y), fabsf(x.z - y.z), fabsf(x.w - y.w)};
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += d[k] < 1.f ? 0.5f * d[k] * d[k] : d[k] - 0.5f;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    rowloss[r] = t * wr;
  }
}
// out[r,f] = d_loss · w[r] · clamp(feat[r,f] − target[r,f], −1, 1)   (SmoothL1' with β = 1); out may alias feat
__global__ void __launch_bounds__(256)
smooth_l1_bwd_kernel(const float* feat, const float* __restrict__ target, const float* __restrict__ w,
                     const float* __restrict__ d_loss, int F, float* out) {
  const int r = blockIdx.x;
  const float g = w[r] * __ldg(d_loss);
  const float4* a = reinterpret_cast<const float4*>(feat + static_cast<size_t>(r) * F);
  const float4* b = reinterpret_cast<const float4*>(target + static_cast<size_t>(r) * F);
  float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(r) * F);
  for (int c = threadIdx.x; c < F / 4; c += blockDim.x) {
    const float4 x = a[c], y = __ldg(b + c);
    o[c] = make_float4(g * fminf(fmaxf(x.x - y.x, -1.f), 1.f), g * fminf(fmaxf(x.y - y.y, -1.f), 1.f),
                       g * fminf(fmaxf(x.z - y.z, -1.f), 1.f), g * fminf(fmaxf(x.w - y.w, -1.f), 1.f));
  }
}
// out[0] = Σ_i x[i] in double, fixed order (single CTA)
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ x, int M, float* out) {
  __shared__ double s_l[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += 1024) acc += static_cast<double>(x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += s_l[i];
    out[0] = static_cast<float>(t);
  }
}
// w[i] = vis_mask[row] ? 1 / (B · max(Σ_v vis_mask[b, v], 1) · F) : 0 with row = rows ? rows[i] : i, b = row / V:
// the per-cell weight of  mean_F → masked sum → ÷ n_mask.clamp(min=1) → batch mean  (modeling.py:278-282)
__global__ void feat_row_weight_kernel(const uint8_t* __restrict__ vis_mask, int B, int V, int F,
                                       const int64_t* __restrict__ rows, int n, float* w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = rows ? rows[i] : i;
  const int64_t b = row / V;
  float v = 0.f;
  if (vis_mask[row]) {
    int cnt = 0;
    for (int j = 0; j < V; ++j) cnt += vis_mask[b * V + j] ? 1 : 0;
    v = 1.0f / (static_cast<float>(B) * static_cast<float>(cnt < 1 ? 1 : cnt) * static_cast<float>(F));
  }
  w[i] = v;
}
// [M, C] fp32 → split [M, Cp], columns C..Cp-1 zero (an upstream logits gradient entering the padded GEMM layout)
__global__ void split_pad_kernel(const float* __restrict__ x, int C, int Cp, bf16* hi, bf16* lo) {
  const size_t r = blockIdx.x;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    bf16 h, l;
    split_bf16(c < C ? x[r * C + c] : 0.f, h, l);
    hi[r * Cp + c] = h;
    if (lo) lo[r * Cp + c] = l;
  }
}

// single CTA, fixed reduction order: deterministic
__global__ void __launch_bounds__(1024)
ce_finish_kernel(const float* __restrict__ rowloss, const int64_t* __restrict__ labels, int64_t ignore, int M,
                 float* stats) {
  __shared__ double s_l[32];
  __shared__ int s_n[32];
  double acc = 0.0;
  int n = 0;
  for (int i = threadIdx.x; i < M; i += 1024)
    if (labels[i] != ignore) { acc += static_cast<double>(rowloss[i]); ++n; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if ((threadIdx.x & 31) == 0) { s_l[threadIdx.x >> 5] = acc; s_n[threadIdx.x >> 5] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int c = 0;
    for (int i = 0; i < 32; ++i) { t += s_l[i]; c += s_n[i]; }
    stats[0] = static_cast<float>(t / static_cast<double>(c));   // 0/0 → NaN like torch when nothing is labelled
    stats[1] = static_cast<float>(c);
  }
}
__global__ void __launch_bounds__(CE_T)
ce_bwd_kernel(const float* __restrict__ logits, int ld, int C, int Cp, const int64_t* __restrict__ labels, int64_t ignore,
              const float* __restrict__ lse, const float* __restrict__ stats, const float* __restrict__ d_loss, bf16* hi,
              bf16* lo) {
  const int m = blockIdx.x;
  const int64_t y = labels[m];
  const float* row = logits + static_cast<size_t>(m) * ld;
  const size_t o = static_cast<size_t>(m) * Cp;
  const float scale = (y == ignore) ? 0.f : __ldg(d_loss) / __ldg(stats + 1);
  const float l = lse[m];
  for (int c = threadIdx.x * 4; c < Cp; c += CE_T * 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = c + j;
      v[j] = (cc < C && scale != 0.f) ? (expf(__ldg(row + cc) - l) - (cc == y ? 1.f : 0.f)) * scale : 0.f;
    }
    store_split4(hi, lo, o + c, make_float4(v[0], v[1], v[2], v[3]));
  }
}
__global__ void __launch_bounds__(CE_T)
softmax_argmax_kernel(const float* __restrict__ logits, int ld, int C, float* prob, int64_t* id) {
  __shared__ float s_v[CE_T / 32];
  __shared__ int s_i[CE_T / 32];
  const int m = blockIdx.x;
  const float* row = logits + static_cast<size_t>(m) * ld;
  float mx, sum;
  row_max_sumexp(row, C, mx, sum);
  // first index whose probability exp(x − max)/Σ is maximal: exp is monotone, so that is the first index with
  // exp(x − max) == 1, i.e. the first x that rounds to the maximum under expf (mirrors torch.max over softmax)
  int best = 0x7fffffff;
  for (int c = threadIdx.x; c < C; c += CE_T)
    if (expf(__ldg(row + c) - mx) == 1.0f) { best = c; break; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_i[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = s_i[0];
    for (int i = 1; i < CE_T / 32; ++i) b = min(b, s_i[i]);
    id[m] = b;
    prob[m] = 1.0f / sum;
  }
  (void)s_v;
}

// warp per row over the partial statistics of a rowstat GEMM epilogue
__global__ void rowstat_merge_kernel(const float* __restrict__ rs, int M, int slots, float* prob, int64_t* id) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* p = rs + static_cast<size_t>(row) * slots * 3;
  float mx = -INFINITY;
  for (int i = lane; i < slots; i += 32) mx = fmaxf(mx, __ldg(p + 3 * i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  int best = 0x7fffffff;
  for (int i = lane; i < slots; i += 32) {
    const float m = __ldg(p + 3 * i), s = __ldg(p + 3 * i + 1);
    const float w = expf(m - mx);                    // 0 for an empty partial (m = −inf)
    sum += s * w;
    if (w == 1.0f) best = min(best, __float_as_int(__ldg(p + 3 * i + 2)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  }
  if (lane == 0) {
    prob[row] = 1.0f / sum;
    id[row] = best;
  }
}

// thread per row, C ≤ 8
__global__ void small_ce_kernel(const float* __restrict__ logits, int M, int C, const int64_t* __restrict__ labels,
                                int64_t ignore, float* rowloss, const float* __restrict__ stats,
                                const float* __restrict__ d_loss, float* dlogits) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* row = logits + static_cast<size_t>(m) * C;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
  const float lse = mx + logf(sum);
  const int64_t y = labels[m];
  if (rowloss) rowloss[m] = (y == ignore) ? 0.f : lse - row[y];
  if (dlogits) {
    const float scale = (y == ignore) ? 0.f : __ldg(d_loss) / __ldg(stats + 1);
    for (int c = 0; c < C; ++c)
      dlogits[static_cast<size_t>(m) * C + c] = scale == 0.f ? 0.f : (expf(row[c] - lse) - (c == y ? 1.f : 0.f)) * scale;
  }
}

// ---- tiny-N linear (matched head) ---------------------------------------------------------------------
// one warp per row
__global__ void small_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                        const float* __restrict__ b, int M, int K, int N, float* y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  for (int n = 0; n < N; ++n) {
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a += __ldg(x + static_cast<size_t>(row) * K + k) * __ldg(W + static_cast<size_t>(n) * K + k);
    a = warp_sum(a);
    if (lane == 0) y[static_cast<size_t>(row) * N + n] = a + __ldg(b + n);
  }
}
// dx[m,k] = Σ_n dy[m,n] W[n,k]
__global__ void small_linear_dx_kernel(const float* __restrict__ dy, const float* __restrict__ W, int M, int K, int N,
                                       float* dx) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(M) * K) return;
  const int m = i / K, k = i % K;
  float a = 0.f;
  for (int n = 0; n < N; ++n) a += __ldg(dy + static_cast<size_t>(m) * N + n) * __ldg(W + static_cast<size_t>(n) * K + k);
  dx[i] = a;
}
// dW[n,k] = Σ_m dy[m,n] x[m,k] (thread per (n,k), fixed order); db[n] = Σ_m dy[m,n]
__global__ void small_linear_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, int M, int K, int N,
                                       float* dW, float* db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * K) {
    const int n = i / K, k = i % K;
    float a = 0.f;
    for (int m = 0; m < M; ++m) a += __ldg(dy + static_cast<size_t>(m) * N + n) * __ldg(x + static_cast<size_t>(m) * K + k);
    dW[i] = a;
  } else if (i < N * K + N) {
    const int n = i - N * K;
    float a = 0.f;
    for (int m = 0; m < M; ++m) a += __ldg(dy + static_cast<size_t>(m) * N + n);
    db[n] = a;
  }
}

// One CTA (256 threads) per 64×64 tile of one job: coalesced fp32 reads, row-major split written directly, the
// transposed split through a padded shared-memory tile.
constexpr int kMaxSplitJobs = 192;
struct SplitJobTable { SplitJob j[kMaxSplitJobs]; int n; };
__global__ void __launch_bounds__(256)
split_batch_kernel(const __grid_constant__ SplitJobTable T) {
  __shared__ float tile[64][65];
  int lo = 0, hi = T.n - 1;                   // job owning this tile: last job with tile0 ≤ blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (T.j[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const SplitJob J = T.j[lo];
  const int t = blockIdx.x - J.tile0;
  const int tiles_c = (J.C + 63) >> 6;
  const int r0 = (t / tiles_c) * 64, c0 = (t % tiles_c) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 column quads × 16 rows per pass
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = r0 + pass * 16 + ty, c = c0 + tx * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < J.R && c < J.C) v = __ldg(reinterpret_cast<const float4*>(J.src + static_cast<size_t>(r) * J.C + c));
    if (J.hi && r < J.R && c < J.C) store_split4(J.hi, J.lo, (static_cast<size_t>(J.row0) + r) * J.C + c, v);
    tile[pass * 16 + ty][tx * 4 + 0] = v.x; tile[pass * 16 + ty][tx * 4 + 1] = v.y;
    tile[pass * 16 + ty][tx * 4 + 2] = v.z; tile[pass * 16 + ty][tx * 4 + 3] = v.w;
  }
  if (!J.t_hi) return;
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int c = c0 + pass * 16 + ty, r = r0 + tx * 4;       // output row = source column
    if (c < J.C && r < J.R) {
      const float4 v = make_float4(tile[tx * 4 + 0][pass * 16 + ty], tile[tx * 4 + 1][pass * 16 + ty],
                                   tile[tx * 4 + 2][pass * 16 + ty], tile[tx * 4 + 3][pass * 16 + ty]);
      store_split4(J.t_hi, J.t_lo, static_cast<size_t>(c) * J.ld_t + J.col0_t + r, v);
    }
  }
}

inline int grid_for(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  return static_cast<int>(b < 148 * 16 ? (b ? b : 1) : 148 * 16);
}

// ---- LayerNorm -------------------------------------------------------------------------------
// One warp per row; a lane owns NV float4 chunks (columns 4·(lane + 32k) …).  H = NV·128.
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              int M, float out_scale, const float* __restrict__ addend, bf16* hi, bf16* lo, float* out_f32,
              float* mean, float* rstd, const DropSite drop) {
  pdl_trigger();
  pdl_wait();
  constexpr int H = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* src = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H);
  float4 x[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    x[k] = __ldg(src + lane + 32 * k);
    s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
  }
  const float mu = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    x[k].x -= mu; x[k].y -= mu; x[k].z -= mu; x[k].w -= mu;
    q += (x[k].x * x[k].x + x[k].y * x[k].y) + (x[k].z * x[k].z + x[k].w * x[k].w);
  }
  const float var = warp_sum(q) * (1.0f / H);
  const float r = 1.0f / sqrtf(var + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = r;
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = (x[k].x * r * g.x + b.x) * out_scale; o.y = (x[k].y * r * g.y + b.y) * out_scale;
    o.z = (x[k].z * r * g.z + b.z) * out_scale; o.w = (x[k].w * r * g.w + b.w) * out_scale;
    const size_t idx = static_cast<size_t>(row) * H + c4 * 4;
    if (addend) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(addend + idx));
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    if (drop.threshold) {     // dropout on the module output (embeddings HF:212, visual feature encoder HF:482)
      const float4 m = drop_hidden4(drop, static_cast<size_t>(row), H, c4 * 4);
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + idx) = o;
    if (hi) store_split4(hi, lo, idx, o);
  }
}

// Warps stride over rows; per-lane dgamma/dbeta partials are reduced across the CTA's warps through
// shared memory and written to part[0/1][blockIdx.x][H].
template <int NV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, float dy_scale, const float* __restrict__ y,
              const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int M,
              float* dx, bf16* dx_hi, bf16* dx_lo, float* part, const DropSite drop_in, const DropSite drop_out,
              bf16* dxm_hi, bf16* dxm_lo) {
  pdl_trigger();
  pdl_wait();
  constexpr int H = NV * 128;
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float4 dg[NV], db[NV], ds[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    dg[k] = make_float4(0, 0, 0, 0); db[k] = make_float4(0, 0, 0, 0); ds[k] = make_float4(0, 0, 0, 0);
  }
  for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
    const float mu = mean[row], r = rstd[row];
    const float4* py = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H);
    const float4* pd = reinterpret_cast<const float4*>(dy + static_cast<size_t>(row) * H);
    float4 xh[NV], dxh[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const float4 yv = __ldg(py + c4);
      float4 d = pd[c4];
      d.x *= dy_scale; d.y *= dy_scale; d.z *= dy_scale; d.w *= dy_scale;
      if (drop_in.threshold) {   // the forward dropped this LayerNorm's OUTPUT: the same mask gates its gradient
        const float4 m = drop_hidden4(drop_in, static_cast<size_t>(row), H, c4 * 4);
        d.x *= m.x; d.y *= m.y; d.z *= m.z; d.w *= m.w;
      }
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      xh[k] = make_float4((yv.x - mu) * r, (yv.y - mu) * r, (yv.z - mu) * r, (yv.w - mu) * r);
      dxh[k] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      dg[k].x += d.x * xh[k].x; dg[k].y += d.y * xh[k].y; dg[k].z += d.z * xh[k].z; dg[k].w += d.w * xh[k].w;
      db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
      c1 += (dxh[k].x + dxh[k].y) + (dxh[k].z + dxh[k].w);
      c2 += (dxh[k].x * xh[k].x + dxh[k].y * xh[k].y) + (dxh[k].z * xh[k].z + dxh[k].w * xh[k].w);
    }
    c1 = warp_sum(c1) * (1.0f / H);
    c2 = warp_sum(c2) * (1.0f / H);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float4 o;
      o.x = r * (dxh[k].x - c1 - xh[k].x * c2); o.y = r * (dxh[k].y - c1 - xh[k].y * c2);
      o.z = r * (dxh[k].z - c1 - xh[k].z * c2); o.w = r * (dxh[k].w - c1 - xh[k].w * c2);
      const size_t idx = static_cast<size_t>(row) * H + (lane + 32 * k) * 4;
      if (dx) *reinterpret_cast<float4*>(dx + idx) = o;
      if (dx_hi) store_split4(dx_hi, dx_lo, idx, o);
      if (drop_out.threshold) {
        // the LayerNorm input was dropout(dense(x)) + residual: the residual path takes dx as is (above), the dense
        // path — its weight / input gradients and its bias gradient (the column sums below) — takes dx ∘ mask
        const float4 m = drop_hidden4(drop_out, static_cast<size_t>(row), H, (lane + 32 * k) * 4);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
        if (dxm_hi) store_split4(dxm_hi, dxm_lo, idx, o);
      }
      ds[k].x += o.x; ds[k].y += o.y; ds[k].z += o.z; ds[k].w += o.w;
    }
  }
  // cross-warp reduction of the column partials: dgamma, dbeta and Σ_rows dx (the bias gradient of the Linear that
  // produced the LayerNorm input)
#pragma unroll
  for (int v = 0; v < 3; ++v) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      __syncthreads();
      red[warp][lane] = v == 0 ? dg[k] : (v == 1 ? db[k] : ds[k]);
      __syncthreads();
      if (warp == 0) {
        float4 a = red[0][lane];
        for (int w = 1; w < nwarp; ++w) {
          a.x += red[w][lane].x; a.y += red[w][lane].y; a.z += red[w][lane].z; a.w += red[w][lane].w;
        }
        float* dst = part + (static_cast<size_t>(v) * gridDim.x + blockIdx.x) * H + (lane + 32 * k) * 4;
        *reinterpret_cast<float4*>(dst) = a;
      }
    }
  }
}

// out_v[h] = Σ_blk part[v][blk][h]; CTA = 32 columns × 8 block lanes, grid (ceil(H/32), nvec); fixed order
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ part, int nblk, int H, float* o0, float* o1, float* o2, float* o3,
                     float* o4, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int h = blockIdx.x * 32 + tx;
  const int v = blockIdx.y;
  float* out = v == 0 ? o0 : v == 1 ? o1 : v == 2 ? o2 : v == 3 ? o3 : o4;
  if (!out) return;
  float a0 = 0.f, a1 = 0.f;
  if (h < H) {
    const float* p = part + static_cast<size_t>(v) * nblk * H + h;
    int b = ty;
    for (; b + 8 < nblk; b += 16) { a0 += __ldg(p + static_cast<size_t>(b) * H); a1 += __ldg(p + static_cast<size_t>(b + 8) * H); }
    if (b < nblk) a0 += __ldg(p + static_cast<size_t>(b) * H);
  }
  red[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && h < H) {
    float sum = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) sum += red[w][tx];
    out[h] = accumulate ? out[h] + sum : sum;
  }
}

// Batched variant: the job table travels as a kernel parameter; CTA → (job, vector, 32-column block)
struct FinishBatch { FinishJob j[64]; int n; };
__global__ void __launch_bounds__(256) colsum_finish_batched_kernel(const __grid_constant__ FinishBatch b) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  int ji = 0;
  while (ji + 1 < b.n && static_cast<int>(blockIdx.x) >= b.j[ji + 1].block0) ++ji;
  const FinishJob& job = b.j[ji];
  const int local = blockIdx.x - job.block0, hb = (job.H + 31) / 32;
  const int v = local / hb, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int h = (local % hb) * 32 + tx, H = job.H, nblk = job.nblk;
  float* out = job.out[v];
  if (!out) return;
  float a0 = 0.f, a1 = 0.f;
  if (h < H) {     // same summation order as colsum_finish_kernel
    const float* p = job.part + static_cast<size_t>(v) * nblk * H + h;
    int k = ty;
    for (; k + 8 < nblk; k += 16) { a0 += __ldg(p + static_cast<size_t>(k) * H); a1 += __ldg(p + static_cast<size_t>(k + 8) * H); }
    if (k < nblk) a0 += __ldg(p + static_cast<size_t>(k) * H);
  }
  red[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && h < H) {
    float sum = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) sum += red[w][tx];
    out[h] = sum;
  }
}

// Column sums: CTA = 32 column quads (128 columns) × 8 row lanes; grid (ceil(N/128), nblk).
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ xf, const bf16* __restrict__ xh, const bf16* __restrict__ xl, int M, int N,
              int ld, float* scratch, const uint8_t* __restrict__ rowmask) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 4;
  float4 a = make_float4(0, 0, 0, 0);
  if (col < N) {
    for (int m = blockIdx.y * 8 + ty; m < M; m += gridDim.y * 8) {
      if (rowmask && !rowmask[m]) continue;
      const size_t idx = static_cast<size_t>(m) * ld + col;
      float4 v = xf ? __ldg(reinterpret_cast<const float4*>(xf + idx)) : load_split4(xh, xl, idx);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && col < N) {
    for (int w = 1; w < 8; ++w) { a.x += red[w][tx].x; a.y += red[w][tx].y; a.z += red[w][tx].z; a.w += red[w][tx].w; }
    *reinterpret_cast<float4*>(scratch + static_cast<size_t>(blockIdx.y) * N + col) = a;
  }
}

// ---- box-position linear ------------------------------------------------------------------------
__global__ void box_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ Wp,
                               const float* __restrict__ bp, int M, int H, float* y2) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(M) * H) return;
  const int m = i / H, h = i % H;
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + m);
  const float4 w = __ldg(reinterpret_cast<const float4*>(Wp) + h);
  // same association as a K=4 dot product followed by the bias add
  y2[i] = (((p.x * w.x + p.y * w.y) + p.z * w.z) + p.w * w.w) + __ldg(bp + h);
}
// CTA = 128 columns × 2 row lanes (256 threads); grid (ceil(H/128), nblk).  part[5][nblk][H]
__global__ void __launch_bounds__(256)
box_bwd_kernel(const float* __restrict__ dy2, const float* __restrict__ pos, int M, int H, float* part) {
  __shared__ float red[5][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int h = blockIdx.x * 128 + tx;
  float a[5] = {0, 0, 0, 0, 0};
  if (h < H) {
    for (int m = blockIdx.y * 2 + ty; m < M; m += gridDim.y * 2) {
      const float d = __ldg(dy2 + static_cast<size_t>(m) * H + h);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + m);
      a[0] += d * p.x; a[1] += d * p.y; a[2] += d * p.z; a[3] += d * p.w; a[4] += d;
    }
  }
  if (ty == 1) for (int v = 0; v < 5; ++v) red[v][tx] = a[v];
  __syncthreads();
  if (ty == 0 && h < H)
    for (int v = 0; v < 5; ++v)
      part[(static_cast<size_t>(v) * gridDim.y + blockIdx.y) * H + h] = a[v] + red[v][tx];
}
// dWp[h][j] = t[j][h]
__global__ void box_pack_kernel(const float* t, int H, float* dWp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * H) dWp[i] = t[(i & 3) * H + (i >> 2)];
}


// Ordered stream compaction of the rows whose label is not `ignore`: one CTA, each thread owns a run of
// consecutive rows, an inclusive scan over the per-thread counts places them.  M ≤ a few 10^4 (B·V or B·L).
__global__ void __launch_bounds__(1024) labelled_rows_kernel(const int64_t* __restrict__ labels, int M,
                                                             int64_t ignore, int64_t* __restrict__ rows,
                                                             int32_t* __restrict__ count) {
  __shared__ int warp_tot[32];
  const int t = threadIdx.x, per = (M + 1023) / 1024;
  const int r0 = min(t * per, M), r1 = min(r0 + per, M);
  int n = 0;
  for (int r = r0; r < r1; ++r) n += labels[r] != ignore;
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += v;
  }
  if ((t & 31) == 31) warp_tot[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    int w = warp_tot[t], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += v;
    }
    warp_tot[t] = wi - w;   // exclusive prefix of the warp totals
    if (t == 31) *count = wi;
  }
  __syncthreads();
  int at = warp_tot[t >> 5] + incl - n;
  for (int r = r0; r < r1; ++r)
    if (labels[r] != ignore) rows[at++] = r;
}

// dst[rows[i], :] = src[i, :]; dst was zero-filled by the caller.
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ src,
                                                           const int64_t* __restrict__ rows, int cols,
                                                           float* __restrict__ dst) {
  const float4* in = reinterpret_cast<const float4*>(src + static_cast<size_t>(blockIdx.x) * cols);
  float4* out = reinterpret_cast<float4*>(dst + static_cast<size_t>(rows[blockIdx.x]) * cols);
  for (int c = threadIdx.x; c < cols / 4; c += blockDim.x) out[c] = in[c];
}
__global__ void gather_i64_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ rows, int n,
                                  int64_t* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[rows[i]];
}
}  // namespace

long long aux_launch_count() { return g_aux_launches.load(); }
int reduce_max_blocks() { return kMaxBlocks; }

int split_f32(const float* x, Split out, size_t n, cudaStream_t s) {
  if (n % 4) return -2;
  if (!n) return 0;
  split_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(x), out.hi, out.lo, n / 4);
  return launch_rc();
}
int unsplit_f32(Split in, float* out, size_t n, cudaStream_t s) {
  if (n % 4) return -2;
  if (!n) return 0;
  unsplit_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(in.hi, in.lo, reinterpret_cast<float4*>(out), n / 4);
  return launch_rc();
}
int fill_f32(float* dst, float v, size_t n, cudaStream_t s) {
  if (!n) return 0;
  fill_kernel<<<grid_for(n, 256), 256, 0, s>>>(dst, v, n);
  return launch_rc();
}
int concat3_f32(const float* a, const float* b, const float* c, float* dst, int n, cudaStream_t s) {
  launch_pdl(concat3_kernel, dim3((3 * n + 255) / 256), dim3(256), 0, s, a, b, c, dst, n);
  return launch_rc();
}
int gather_rows(const float* table, const int64_t* ids, const uint8_t* mask, const float* fill, int rows, int cols,
                float* out_f32, Split out, cudaStream_t s) {
  if (cols % 4) return -2;
  if (!rows) return 0;
  gather_rows_kernel<<<rows, 256, 0, s>>>(table, ids, mask, fill, cols, out_f32, out.hi, out.lo);
  return launch_rc();
}

int labelled_rows(const int64_t* labels, int M, int64_t ignore, int64_t* rows, int32_t* count, cudaStream_t s) {
  if (M < 1) return -2;
  labelled_rows_kernel<<<1, 1024, 0, s>>>(labels, M, ignore, rows, count);
  return launch_rc();
}
int scatter_rows(const float* src, const int64_t* rows, int n, int M, int cols, float* dst, cudaStream_t s) {
  if (cols % 4) return -2;
  cudaError_t e = cudaMemsetAsync(dst, 0, static_cast<size_t>(M) * cols * sizeof(float), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (!n) return 0;
  scatter_rows_kernel<<<n, 256, 0, s>>>(src, rows, cols, dst);
  return launch_rc();
}
int gather_i64(const int64_t* src, const int64_t* rows, int n, int64_t* dst, cudaStream_t s) {
  if (!n) return 0;
  gather_i64_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, rows, n, dst);
  return launch_rc();
}

int split_batch(SplitJob* jobs, int njobs, cudaStream_t s) {
  for (int first = 0; first < njobs; first += kMaxSplitJobs) {
    SplitJobTable T;
    T.n = njobs - first < kMaxSplitJobs ? njobs - first : kMaxSplitJobs;
    int tiles = 0;
    for (int i = 0; i < T.n; ++i) {
      SplitJob j = jobs[first + i];
      if ((j.C % 4) || (j.R % 4) || (j.ld_t % 4) || (j.col0_t % 4)) return -2;
      j.tile0 = tiles;
      tiles += ((j.R + 63) / 64) * ((j.C + 63) / 64);
      T.j[i] = j;
    }
    split_batch_kernel<<<tiles, 256, 0, s>>>(T);
    int rc = launch_rc();
    if (rc) return rc;
  }
  return 0;
}
int split_rows_f32(const float* x, size_t ld_in, int rows, int cols, Split out, cudaStream_t s) {
  if ((cols % 4) || (ld_in % 4)) return -2;
  if (!rows) return 0;
  split_rows_kernel<<<rows, 192, 0, s>>>(x, ld_in, cols, out.hi, out.lo);
  return launch_rc();
}
int tanh_bwd_split(const float* d, const float* y, Split out, size_t n, cudaStream_t s) {
  if (n % 4) return -2;
  if (!n) return 0;
  tanh_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(d),
                                                        reinterpret_cast<const float4*>(y), out.hi, out.lo, n / 4);
  return launch_rc();
}
int embed_sum(const int64_t* ids, const int64_t* tt, const float* word, const float* pos, const float* type, int rows,
              int L, int H, float* y, cudaStream_t s) {
  if (H % 4) return -2;
  if (!rows) return 0;
  embed_sum_kernel<<<rows, 192, 0, s>>>(ids, tt, word, pos, type, L, H, y);
  return launch_rc();
}
int embed_scatter(const int64_t* ids, const int64_t* tt, const float* dy, int rows, int L, int H, float* dword,
                  float* dpos, float* dtype, cudaStream_t s) {
  if (!rows) return 0;
  embed_scatter_kernel<<<rows, 256, 0, s>>>(ids, tt, dy, L, H, dword, dpos, dtype);
  return launch_rc();
}

int gelu_bwd_split(const float* dg, const float* u, Split out, size_t n, cudaStream_t s) {
  if (n % 4) return -2;
  if (!n) return 0;
  gelu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(dg),
                                                        reinterpret_cast<const float4*>(u), out.hi, out.lo, n / 4);
  return launch_rc();
}
int ce_fwd(const float* logits, int ld, int M, int C, const int64_t* labels, int64_t ignore_index, float* lse,
           float* rowloss, float* stats, cudaStream_t s) {
  if (ld % 4) return -2;
  if (M < 1 || C < 1) return -21;
  ce_fwd_kernel<<<M, CE_T, 0, s>>>(logits, ld, C, labels, ignore_index, lse, rowloss);
  int rc = launch_rc();
  if (rc) return rc;
  ce_finish_kernel<<<1, 1024, 0, s>>>(rowloss, labels, ignore_index, M, stats);
  return launch_rc();
}
int ce_bwd(const float* logits, int ld, int M, int C, int Cp, const int64_t* labels, int64_t ignore_index,
           const float* lse, const float* stats, const float* d_loss, Split dlogits, cudaStream_t s) {
  if ((ld % 4) || (Cp % 4) || Cp < C) return -2;
  if (M < 1) return -21;
  ce_bwd_kernel<<<M, CE_T, 0, s>>>(logits, ld, C, Cp, labels, ignore_index, lse, stats, d_loss, dlogits.hi, dlogits.lo);
  return launch_rc();
}
int smooth_l1_fwd(const float* feat, const float* target, const float* w, int M, int F, float* rowloss, float* loss,
                  cudaStream_t s) {
  if (F % 4) return -2;
  if (M < 1) return -21;
  smooth_l1_fwd_kernel<<<M, 256, 0, s>>>(feat, target, w, F, rowloss);
  int rc = launch_rc();
  if (rc) return rc;
  sum_rows_kernel<<<1, 1024, 0, s>>>(rowloss, M, loss);
  return launch_rc();
}
int smooth_l1_bwd(const float* feat, const float* target, const float* w, const float* d_loss, int M, int F, float* out,
                  cudaStream_t s) {
  if (F % 4) return -2;
  if (M < 1) return -21;
  smooth_l1_bwd_kernel<<<M, 256, 0, s>>>(feat, target, w, d_loss, F, out);
  return launch_rc();
}
int feat_row_weight(const uint8_t* vis_mask, int B, int V, int F, const int64_t* rows, int n, float* w, cudaStream_t s) {
  if (B < 1 || V < 1 || F < 1) return -21;
  if (!n) return 0;
  feat_row_weight_kernel<<<(n + 127) / 128, 128, 0, s>>>(vis_mask, B, V, F, rows, n, w);
  return launch_rc();
}
int split_pad_f32(const float* x, int M, int C, int Cp, Split out, cudaStream_t s) {
  if (Cp < C || M < 1) return -21;
  split_pad_kernel<<<M, 256, 0, s>>>(x, C, Cp, out.hi, out.lo);
  return launch_rc();
}
int softmax_argmax(const float* logits, int ld, int M, int C, float* prob, int64_t* id, cudaStream_t s) {
  if (ld % 4) return -2;
  if (M < 1 || C < 1) return -21;
  softmax_argmax_kernel<<<M, CE_T, 0, s>>>(logits, ld, C, prob, id);
  return launch_rc();
}
int rowstat_merge(const float* rowstat, int M, int slots, float* prob, int64_t* id, cudaStream_t s) {
  if (M < 1 || slots < 1) return -21;
  launch_pdl(rowstat_merge_kernel, dim3((M + 7) / 8), dim3(256), 0, s, rowstat, M, slots, prob, id);
  return launch_rc();
}
int small_ce_fwd(const float* logits, int M, int C, const int64_t* labels, int64_t ignore_index, float* rowloss,
                 float* stats, cudaStream_t s) {
  if (C < 1 || C > 8) return -1;
  if (M < 1) return -21;
  small_ce_kernel<<<(M + 127) / 128, 128, 0, s>>>(logits, M, C, labels, ignore_index, rowloss, nullptr, nullptr, nullptr);
  int rc = launch_rc();
  if (rc) return rc;
  ce_finish_kernel<<<1, 1024, 0, s>>>(rowloss, labels, ignore_index, M, stats);
  return launch_rc();
}
int small_ce_bwd(const float* logits, int M, int C, const int64_t* labels, int64_t ignore_index, const float* stats,
                 const float* d_loss, float* dlogits, cudaStream_t s) {
  if (C < 1 || C > 8) return -1;
  if (M < 1) return -21;
  small_ce_kernel<<<(M + 127) / 128, 128, 0, s>>>(logits, M, C, labels, ignore_index, nullptr, stats, d_loss, dlogits);
  return launch_rc();
}
int small_linear_fwd(const float* x, const float* W, const float* b, int M, int K, int N, float* y, cudaStream_t s) {
  if (N < 1 || N > 8) return -1;
  if (!M) return 0;
  small_linear_fwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, W, b, M, K, N, y);
  return launch_rc();
}
int small_linear_bwd(const float* dy, const float* x, const float* W, int M, int K, int N, float* dW, float* db,
                     float* dx, cudaStream_t s) {
  if (N < 1 || N > 8) return -1;
  if (!M) return 0;
  small_linear_dw_kernel<<<(N * K + N + 255) / 256, 256, 0, s>>>(dy, x, M, K, N, dW, db);
  int rc = launch_rc();
  if (rc) return rc;
  const size_t n = static_cast<size_t>(M) * K;
  small_linear_dx_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(dy, W, M, K, N, dx);
  return launch_rc();
}

#define XLX_LN_DISPATCH(KERNEL, H, GRID, ...)                                          \
  switch (H) {                                                                         \
    case 128: launch_pdl(KERNEL<1>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break;  \
    case 256: launch_pdl(KERNEL<2>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break;  \
    case 512: launch_pdl(KERNEL<4>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break;  \
    case 768: launch_pdl(KERNEL<6>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break;  \
    case 1024: launch_pdl(KERNEL<8>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break; \
    case 1536: launch_pdl(KERNEL<12>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break; \
    case 2048: launch_pdl(KERNEL<16>, dim3(GRID), dim3(256), 0, s, __VA_ARGS__); break; \
    default: return -4;                                                                \
  }

int layernorm_fwd(const float* y, const float* gamma, const float* beta, float eps, int M, int H, float out_scale,
                  const float* addend, Split out, float* out_f32, float* mean, float* rstd, cudaStream_t s,
                  DropSite drop) {
  if (!M) return 0;
  const int grid = (M + 7) / 8;
  XLX_LN_DISPATCH(ln_fwd_kernel, H, grid, y, gamma, beta, eps, M, out_scale, addend, out.hi, out.lo, out_f32, mean,
                  rstd, drop);
  return launch_rc();
}
int layernorm_bwd(const float* dy, float dy_scale, const float* y, const float* gamma, const float* mean,
                  const float* rstd, int M, int H, float* dx, Split dx_split, float* part, int* nblk_out,
                  cudaStream_t s, DropSite drop_in, DropSite drop_out, Split dx_masked) {
  int grid = (M + 7) / 8;
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  if (grid < 1) grid = 1;
  *nblk_out = grid;
  if (drop_out.threshold && !dx_masked.hi) return -24;
  XLX_LN_DISPATCH(ln_bwd_kernel, H, grid, dy, dy_scale, y, gamma, mean, rstd, M, dx, dx_split.hi, dx_split.lo, part,
                  drop_in, drop_out, dx_masked.hi, dx_masked.lo);
  return launch_rc();
}

// ---- materialised dropout masks (test / debugging aid: the product kernels regenerate them on the fly) -----------
__global__ void drop_mask_hidden_kernel(const DropSite d, size_t rows, int H, float* out) {
  const size_t n4 = rows * H / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    reinterpret_cast<float4*>(out)[i] = drop_hidden4(d, e / H, H, static_cast<int>(e % H));
  }
}
__global__ void drop_mask_probs_kernel(const DropSite d, size_t rows, int Sk, float* out) {
  const int pairs = (Sk + 1) >> 1;
  const size_t n = rows * pairs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / pairs;
    const int j = static_cast<int>(i % pairs) * 2;
    const float2 m = drop_prob2(d, r, Sk, j);
    out[r * Sk + j] = m.x;
    if (j + 1 < Sk) out[r * Sk + j + 1] = m.y;
  }
}
int dropout_mask_hidden(DropSite d, size_t rows, int H, float* out, cudaStream_t s) {
  if (!rows || H % 4) return -2;
  drop_mask_hidden_kernel<<<grid_for(rows * H / 4, 256), 256, 0, s>>>(d, rows, H, out);
  return launch_rc();
}
int dropout_mask_probs(DropSite d, size_t rows, int Sk, float* out, cudaStream_t s) {
  if (!rows || Sk < 1) return -21;
  drop_mask_probs_kernel<<<grid_for(rows * ((Sk + 1) / 2), 256), 256, 0, s>>>(d, rows, Sk, out);
  return launch_rc();
}
int colsum_finish(const float* part, int nvec, int nblk, int H, float* const* outs, int accumulate, cudaStream_t s) {
  if (nvec < 1 || nvec > 5) return -1;
  float* o[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int v = 0; v < nvec; ++v) o[v] = outs[v];
  launch_pdl(colsum_finish_kernel, dim3((H + 31) / 32, nvec), dim3(256), 0, s, part, nblk, H, o[0], o[1], o[2], o[3],
             o[4], accumulate);
  return launch_rc();
}
int colsum_finish_batched(FinishJob* jobs, int njobs, cudaStream_t s) {
  for (int j0 = 0; j0 < njobs; j0 += 64) {
    FinishBatch b;
    b.n = njobs - j0 < 64 ? njobs - j0 : 64;
    int blocks = 0;
    for (int i = 0; i < b.n; ++i) {
      FinishJob j = jobs[j0 + i];
      if (j.nvec < 1 || j.nvec > 3 || j.nblk < 1 || j.H < 1) return -1;
      j.block0 = blocks;
      blocks += ((j.H + 31) / 32) * j.nvec;
      b.j[i] = j;
    }
    launch_pdl(colsum_finish_batched_kernel, dim3(blocks), dim3(256), 0, s, b);
    int rc = launch_rc();
    if (rc) return rc;
  }
  return 0;
}
int colsum_partial(const float* x_f32, Split x, int M, int N, int ld, float* part, int* nblk_out, cudaStream_t s,
                   const uint8_t* rowmask) {
  if ((N % 4) || (ld % 4)) return -2;
  int nblk = (M + 63) / 64;
  if (nblk > 128) nblk = 128;
  if (nblk < 1) nblk = 1;
  *nblk_out = nblk;
  launch_pdl(colsum_kernel, dim3((N + 127) / 128, nblk), dim3(256), 0, s, x_f32, x.hi, x.lo, M, N, ld, part, rowmask);
  return launch_rc();
}
int colsum(const float* x_f32, Split x, int M, int N, int ld, float* scratch, float* out, cudaStream_t s,
           const uint8_t* rowmask) {
  if ((N % 4) || (ld % 4)) return -2;
  int nblk = (M + 63) / 64;
  if (nblk > 128) nblk = 128;
  if (nblk < 1) nblk = 1;
  launch_pdl(colsum_kernel, dim3((N + 127) / 128, nblk), dim3(256), 0, s, x_f32, x.hi, x.lo, M, N, ld, scratch,
             rowmask);
  int rc = launch_rc();
  if (rc) return rc;
  float* outs[1] = {out};
  return colsum_finish(scratch, 1, nblk, N, outs, 0, s);
}

int box_linear_fwd(const float* pos, const float* Wp, const float* bp, int M, int H, float* y2, cudaStream_t s) {
  const size_t n = static_cast<size_t>(M) * H;
  if (!n) return 0;
  box_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(pos, Wp, bp, M, H, y2);
  return launch_rc();
}
int box_linear_bwd(const float* dy2, const float* pos, int M, int H, float* scratch, float* dWp, float* dbp,
                   cudaStream_t s) {
  int nblk = (M + 31) / 32;
  if (nblk > 128) nblk = 128;
  if (nblk < 1) nblk = 1;
  box_bwd_kernel<<<dim3((H + 127) / 128, nblk), 256, 0, s>>>(dy2, pos, M, H, scratch);
  int rc = launch_rc();
  if (rc) return rc;
  // reduce into a [4][H] temp placed after the partials, then interleave into dWp [H][4]
  float* t = scratch + static_cast<size_t>(5) * nblk * H;
  float* outs[5] = {t, t + H, t + 2 * H, t + 3 * H, dbp};
  if ((rc = colsum_finish(scratch, 5, nblk, H, outs, 0, s))) return rc;
  box_pack_kernel<<<(4 * H + 255) / 256, 256, 0, s>>>(t, H, dWp);
  return launch_rc();
}

void count_aux_launch() { g_aux_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace xlx
