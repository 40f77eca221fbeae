// Fused optimiser step of the pre-training loop (x-lxmert/src/pretrain/lxmert_pretrain.py:343-364, :110-141):
// global-norm gradient clipping (torch.nn.utils.clip_grad_norm_) + HF AdamW (transformers 4.1.1
// optimization.AdamW: correct_bias, eps inside the square root's sum, decoupled weight decay applied AFTER the Adam
// update with the plain learning rate) over every parameter tensor in two launches.  HBM-bound: it reads g, p, m, v
// and writes p, m, v once (28 bytes per parameter); the clip coefficient is applied on the fly, gradients are not
// rewritten.  Tensors travel as a kernel-parameter table (no device-side pointer arrays, no copies).
#include <atomic>

#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

using namespace xlx;

namespace {

constexpr int kMaxTensors = 320;          // 320 × 48 B + header < 32 KB of kernel parameters
constexpr int kChunk = 4096;              // elements per CTA work item (256 threads × 4 × 4)
struct TensorRef { float* p; const float* g; float* m; float* v; long long n; float wd; int tile0; };
struct Table { TensorRef t[kMaxTensors]; int n; };

__device__ __forceinline__ int find_tensor(const Table& T, int tile) {
  int lo = 0, hi = T.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (T.t[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// partial[tile] = Σ g² over the tile (fp32 lanes, fixed order): deterministic two-stage norm
__global__ void __launch_bounds__(256)
sqnorm_kernel(const __grid_constant__ Table T, int tile_base, float* partial) {
  __shared__ float red[8];
  const int tile = blockIdx.x;
  const TensorRef R = T.t[find_tensor(T, tile)];
  const long long e0 = static_cast<long long>(tile - R.tile0) * kChunk;
  float acc = 0.f;
  for (int i = threadIdx.x * 4; i < kChunk; i += 256 * 4) {
    const long long e = e0 + i;
    if (e + 4 <= R.n) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(R.g + e));
      acc += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    } else {
      for (int j = 0; j < 4; ++j) if (e + j < R.n) { const float g = R.g[e + j]; acc += g * g; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[tile_base + tile] = s;
  }
}
// out[0] = Σ partial (double accumulation, single CTA, fixed order); accumulate = add to the existing value
__global__ void __launch_bounds__(1024)
sqnorm_finish_kernel(const float* __restrict__ partial, int n, float* out, int accumulate) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += static_cast<double>(partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = accumulate ? static_cast<double>(out[0]) : 0.0;
    for (int i = 0; i < 32; ++i) s += red[i];
    out[0] = static_cast<float>(s);
  }
}

struct Hyper { float lr, beta1, beta2, omb1, omb2, eps, step_size, max_norm; };   // omb = 1 − beta, rounded from double

__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ Table T, const Hyper h, const float* __restrict__ sqnorm) {
  const int tile = blockIdx.x;
  const TensorRef R = T.t[find_tensor(T, tile)];
  const long long e0 = static_cast<long long>(tile - R.tile0) * kChunk;
  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), applied only when < 1
  float coef = 1.f;
  if (sqnorm && h.max_norm > 0.f) {
    const float c = h.max_norm / (sqrtf(__ldg(sqnorm)) + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  const float decay = -h.lr * R.wd;          // p ← p + (−lr·wd)·p after the Adam update (HF AdamW)
  for (int i = threadIdx.x * 4; i < kChunk; i += 256 * 4) {
    const long long e = e0 + i;
    if (e >= R.n) break;
    if (e + 4 <= R.n) {
      float4 g = __ldg(reinterpret_cast<const float4*>(R.g + e));
      float4 p = *reinterpret_cast<float4*>(R.p + e);
      float4 m = *reinterpret_cast<float4*>(R.m + e);
      float4 v = *reinterpret_cast<float4*>(R.v + e);
      float* gp = &g.x; float* pp = &p.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = gp[j] * coef;
        mp[j] = mp[j] * h.beta1 + h.omb1 * gj;
        vp[j] = vp[j] * h.beta2 + h.omb2 * gj * gj;
        pp[j] = pp[j] - h.step_size * (mp[j] / (sqrtf(vp[j]) + h.eps));
        if (R.wd > 0.f) pp[j] = pp[j] + decay * pp[j];
      }
      *reinterpret_cast<float4*>(R.p + e) = p;
      *reinterpret_cast<float4*>(R.m + e) = m;
      *reinterpret_cast<float4*>(R.v + e) = v;
    } else {
      for (int j = 0; j < 4 && e + j < R.n; ++j) {
        const float gj = R.g[e + j] * coef;
        const float m = R.m[e + j] * h.beta1 + h.omb1 * gj;
        const float v = R.v[e + j] * h.beta2 + h.omb2 * gj * gj;
        float p = R.p[e + j] - h.step_size * (m / (sqrtf(v) + h.eps));
        if (R.wd > 0.f) p = p + decay * p;
        R.p[e + j] = p; R.m[e + j] = m; R.v[e + j] = v;
      }
    }
  }
}

std::atomic<long long> g_optim_launches{0};

// fill a table from tensors [first, first + count); returns the tile count
int fill(Table& T, float* const* params, const float* const* grads, float* const* m, float* const* v,
         const int64_t* elems, const float* wd, int first, int count) {
  int tiles = 0;
  T.n = count;
  for (int i = 0; i < count; ++i) {
    TensorRef& R = T.t[i];
    const int k = first + i;
    R.p = params ? params[k] : nullptr; R.g = grads[k]; R.m = m ? m[k] : nullptr; R.v = v ? v[k] : nullptr;
    R.n = elems[k]; R.wd = wd ? wd[k] : 0.f; R.tile0 = tiles;
    tiles += static_cast<int>((elems[k] + kChunk - 1) / kChunk);
  }
  return tiles;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int64_t xlx_optim_launch_count(void) { return g_optim_launches.load(); }

int64_t xlx_optim_scratch_floats(const int64_t* elems, int32_t n) {
  if (!elems || n < 0) return -24;
  int64_t tiles = 0;
  for (int i = 0; i < n; ++i) tiles += (elems[i] + kChunk - 1) / kChunk;
  return tiles + 8;
}

int32_t xlx_grad_sqnorm(const float* const* grads, const int64_t* elems, int32_t n, float* scratch, float* out,
                        void* stream) {
  if (n < 1) return -21;
  if (!grads || !elems || !scratch || !out) return -24;
  XLX_TRY(ensure_device(out));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int base = 0;
  for (int first = 0; first < n; first += kMaxTensors) {
    Table T;
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    for (int i = 0; i < count; ++i)
      if (!grads[first + i] || elems[first + i] < 1 || !aligned16(grads[first + i])) return -2;
    const int tiles = fill(T, nullptr, grads, nullptr, nullptr, elems, nullptr, first, count);
    sqnorm_kernel<<<tiles, 256, 0, st>>>(T, base, scratch);
    g_optim_launches.fetch_add(1, std::memory_order_relaxed);
    XLX_CUDA(cudaGetLastError());
    base += tiles;
  }
  sqnorm_finish_kernel<<<1, 1024, 0, st>>>(scratch, base, out, 0);
  g_optim_launches.fetch_add(1, std::memory_order_relaxed);
  XLX_CUDA(cudaGetLastError());
  return 0;
}

int32_t xlx_adamw_step(float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* elems, const float* weight_decay, int32_t n, double lr,
                       double beta1, double beta2, double eps, int32_t step, int32_t correct_bias, const float* sqnorm,
                       double max_grad_norm, void* stream) {
  if (n < 1 || step < 1) return -21;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !elems) return -24;
  XLX_TRY(ensure_device(params[0]));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Hyper h;
  // scalars are formed in double like the reference's Python arithmetic and rounded to fp32 once
  h.lr = static_cast<float>(lr); h.beta1 = static_cast<float>(beta1); h.beta2 = static_cast<float>(beta2);
  h.omb1 = static_cast<float>(1.0 - beta1); h.omb2 = static_cast<float>(1.0 - beta2);
  h.eps = static_cast<float>(eps); h.max_norm = static_cast<float>(max_grad_norm);
  double step_size = lr;
  if (correct_bias) step_size = lr * sqrt(1.0 - pow(beta2, step)) / (1.0 - pow(beta1, step));
  h.step_size = static_cast<float>(step_size);
  for (int first = 0; first < n; first += kMaxTensors) {
    Table T;
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    for (int i = 0; i < count; ++i) {
      const int k = first + i;
      if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || elems[k] < 1) return -24;
      if (!aligned16(params[k]) || !aligned16(grads[k]) || !aligned16(exp_avg[k]) || !aligned16(exp_avg_sq[k])) return -2;
    }
    const int tiles = fill(T, params, grads, exp_avg, exp_avg_sq, elems, weight_decay, first, count);
    adamw_kernel<<<tiles, 256, 0, st>>>(T, h, sqnorm);
    g_optim_launches.fetch_add(1, std::memory_order_relaxed);
    XLX_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
