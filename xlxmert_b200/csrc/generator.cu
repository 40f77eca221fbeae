// Grid-feature GAN generator forward (image_generator/src/layers.py:135-253, canonical configuration of
// scripts/train_generator.bash + x-lxmert/src/tasks/sample_images.py:53-67: base_dim 32, emb_dim 2048,
// codebook_dim 256, 8×8 → 256×256, SPADE + InstanceNorm, spectral norm) on sm_100a.
//
// Layout: every activation is NHWC ("pixel-major": [B, H, W, C], C innermost) so that a convolution is an implicit
// GEMM whose A operand the TMA unit reads straight from the activation tensor, one shifted box per filter tap
// (gemm_sm100.cuh, ConvGeometry) — there is no im2col buffer.  Convolution inputs are stored as split-bf16 pairs,
// their fp32 outputs feed the normalisation kernels below.  The 128-channel SPADE hidden map never exists in fp32.
//
// Kernels in this file are the memory-bound glue between the tensor-core convolutions: instance-norm statistics,
// SPADE modulation + LeakyReLU (+ fused bilinear ×2), bilinear resampling of the style map, ToRGB accumulation.
#include <math.h>

#include <atomic>

#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"
#include "xlx_ptx.cuh"

using namespace xlx;

namespace {

constexpr int CH = 32;          // width of every block (resolution_channels = min(·, base_dim), layers.py:161-175)
constexpr int RES_PACK = 4;     // pixels per GEMM row of the 1×1 shortcut convolution (see Prep::res)
constexpr int HID = 128;        // SPADE hidden width (layers.py:23)
constexpr int EMB = 2048, CODE = 256, NBLK = 5, R0 = 8, RGB_LD = 16;
constexpr int N_PARAMS = 10 + NBLK * 26 + NBLK * 2;

std::atomic<long long> g_gen_launches{0};
inline int krc() {
  g_gen_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, size_t idx, float4 v) {
  uint2 h, l;
  split_bf16x2(v.x, v.y, h.x, l.x);       // these kernels are instruction-bound, not byte-bound
  split_bf16x2(v.z, v.w, h.y, l.y);
  *reinterpret_cast<uint2*>(hi + idx) = h;
  *reinterpret_cast<uint2*>(lo + idx) = l;
}

// PyTorch upsample_bilinear2d, align_corners=False: source index and weights for output index o.
__device__ __forceinline__ void bilinear_src(int o, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (static_cast<float>(o) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
  l0 = 1.f - l1;
}

// ---- weight preparation ---------------------------------------------------------------------------
// 1/σ of a spectrally normalised conv in eval mode: σ = uᵀ·(W_mat·v) with the stored u, v (torch.nn.utils
// spectral_norm, compute_weight with do_power_iteration=False).  One CTA.
__global__ void sigma_inv_kernel(const float* __restrict__ w, const float* __restrict__ u, const float* __restrict__ v,
                                 int rows, int cols, float* out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int r = threadIdx.x >> 5; r < rows; r += blockDim.x >> 5) {
    float d = 0.f;
    for (int c = threadIdx.x & 31; c < cols; c += 32) d += w[static_cast<size_t>(r) * cols + c] * v[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0) acc += u[r] * d;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
    *out = 1.0f / s;
  }
}
// w [Co, Ci, kh, kw] (torch conv layout) → dst[(row0 + co) * ld + tap * ctot + col0 + ci], tap = ky * kw + kx,
// scaled by *scale (nullable), as split bf16.  For grouped convs col0 is chosen per output channel:
// col0 = (co / co_per_group) * Ci.
__global__ void prep_conv_kernel(const float* __restrict__ w, const float* __restrict__ scale, int Co, int Ci, int taps,
                                 int ctot, int co_per_group, int row0, int ld, bf16* hi, bf16* lo, int col_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * taps) return;
  const int tap = i % taps, ci = (i / taps) % Ci, co = i / (taps * Ci);
  const int col0 = col_base + (co_per_group ? (co / co_per_group) * Ci : 0);
  const float v = w[i] * (scale ? *scale : 1.0f);
  bf16 h, l;
  split_bf16(v, h, l);
  const size_t d = static_cast<size_t>(row0 + co) * ld + static_cast<size_t>(tap) * ctot + col0 + ci;
  hi[d] = h; lo[d] = l;
}
// w [HID, CH, 3, 3] → dst[(tap·HID + co)·CH + ci] as split bf16 (tap-major rows: one GEMM gives all nine W_t·y)
__global__ void prep_tapmajor_kernel(const float* __restrict__ w, int Co, int Ci, bf16* hi, bf16* lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * 9) return;
  const int tap = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
  bf16 h, l;
  split_bf16(w[i], h, l);
  const size_t d = (static_cast<size_t>(tap) * Co + co) * Ci + ci;
  hi[d] = h; lo[d] = l;
}
__global__ void repeat_bias_kernel(const float* b, int n, float* dst, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dst[i] = b[i % n];
}
__global__ void concat_bias_kernel(const float* a, int na, const float* b, int nb, float* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = i < na ? a[i] : (i < na + nb ? b[i - na] : 0.f);
}

// ---- instance-norm statistics -----------------------------------------------------------------------
// x: [B, HW, ld] fp32 (first CH channels used).  acc[b][c][0..1] += (Σx, Σx²) in fp64.  grid (chunks, B), 256 thr.
__global__ void __launch_bounds__(256)
in_stats_kernel(const float* __restrict__ x, int ld, int HW, double* acc) {
  // thread = (pixel lane, channel quad): 16-byte loads, two pixels in flight per thread
  __shared__ double s1[32][CH + 1], s2[32][CH + 1];
  const int c4 = threadIdx.x & 7, lane = threadIdx.x >> 3, b = blockIdx.y;
  const float* xb = x + static_cast<size_t>(b) * HW * ld + c4 * 4;
  double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
  const int step = gridDim.x * 32;
  int p = blockIdx.x * 32 + lane;
  for (; p + step < HW; p += 2 * step) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(p) * ld));
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(p + step) * ld));
    const float uf[4] = {u.x, u.y, u.z, u.w}, vf[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double du = uf[k], dv = vf[k];
      a1[k] += du; a2[k] += du * du;
      a1[k] += dv; a2[k] += dv * dv;
    }
  }
  for (; p < HW; p += step) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(p) * ld));
    const float uf[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const double du = uf[k]; a1[k] += du; a2[k] += du * du; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { s1[lane][c4 * 4 + k] = a1[k]; s2[lane][c4 * 4 + k] = a2[k]; }
  __syncthreads();
  if (threadIdx.x < CH) {
    const int c = threadIdx.x;
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < 32; ++i) { t1 += s1[i][c]; t2 += s2[i][c]; }
    atomicAdd(acc + (static_cast<size_t>(b) * CH + c) * 2, t1);
    atomicAdd(acc + (static_cast<size_t>(b) * CH + c) * 2 + 1, t2);
  }
}
// InstanceNorm2d(affine=False): biased variance over H×W, eps 1e-5 (layers.py:16)
__global__ void in_finish_kernel(const double* __restrict__ acc, int n, int HW, float* mean, float* rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = acc[2 * i] / HW;
  double var = acc[2 * i + 1] / HW - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = static_cast<float>(m);
  rstd[i] = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

// ---- resampling / modulation --------------------------------------------------------------------------
// Bilinear resize of a CH-channel NHWC map [B, Hi, Hi, ld] → split [B, Ho, Ho, CH].  One thread = 4 channels of a pixel.
__global__ void resize_split_kernel(const float* __restrict__ x, int ld, int Hi, int Ho, size_t npix, bf16* hi, bf16* lo) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t >= npix * (CH / 4)) return;
  const int c4 = static_cast<int>(t & 7);             // CH/4 = 8 and Ho is a power of two: no 64-bit divisions
  const size_t p = t >> 3;
  const int lg = 31 - __clz(Ho);
  const int ox = static_cast<int>(p & (Ho - 1)), oy = static_cast<int>((p >> lg) & (Ho - 1));
  const size_t b = p >> (2 * lg);
  const float scale = static_cast<float>(Hi) / static_cast<float>(Ho);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(oy, scale, Hi, y0, y1, ly0, ly1);
  bilinear_src(ox, scale, Hi, x0, x1, lx0, lx1);
  const float* xb = x + b * Hi * Hi * ld + c4 * 4;
  const float4 v00 = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(y0) * Hi + x0) * ld));
  const float4 v01 = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(y0) * Hi + x1) * ld));
  const float4 v10 = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(y1) * Hi + x0) * ld));
  const float4 v11 = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(y1) * Hi + x1) * ld));
  float4 o;
  o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
  o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
  o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
  o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  store_split(hi, lo, p * CH + c4 * 4, o);
}

// ---- SPADE hidden map without its convolution --------------------------------------------------------------
// layers.py:33-47: actv = ReLU(conv3×3(bilinear(y → R×R))) with y the 8×8×32 style map.  Both the resize and the
// convolution are LINEAR in y, so the 32 → 128 convolution never has to run at R×R (73.7 kFLOP per pixel, 2.8 ms of the
// B = 128 forward at 256² alone): with Z_t = W_t·y computed once at 8×8 for the nine taps t (one small tcgen05 GEMM),
//     u(p) = b + Σ_t bilinear(Z_t)(p + t)            (zero outside the image = the conv padding)
// and the bilinear sample separates into a row stage and a column stage:
//     Ry[b, py, tx, c]   = Σ_ty Σ_a  wy_a(py + ty) · Z_(ty,tx)[b, ya(py + ty), c]          (≤ 6 terms; per image ROW)
//     u(py, px)          = b + Σ_tx Σ_a  wx_a(px + tx) · Ry[b, py, tx, xa(px + tx)]          (≤ 6 terms; per pixel)
// — exact (same products, different summation order), 768 FMA per pixel instead of 36 864, HBM-bound on writing the
// split-bf16 hidden map the γ/β convolution reads.
// Z: [B·64, 9·HID] fp32, column (tap·HID + ch), tap = ty·3 + tx;  Ry: [B, R, 3, 8, HID] fp32.
__global__ void __launch_bounds__(256)
spade_rows_kernel(const float* __restrict__ Z, int R, size_t n, float* __restrict__ Ry) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t >= n) return;                                   // n = B·R·3·8·(HID/4)
  const int c4 = static_cast<int>(t & (HID / 4 - 1));
  size_t q = t >> 5;
  const int c = static_cast<int>(q & 7); q >>= 3;
  const int tx = static_cast<int>(q % 3); q /= 3;
  const int lg = 31 - __clz(R);
  const int py = static_cast<int>(q & (R - 1));
  const size_t b = q >> lg;
  const float scale = static_cast<float>(R0) / static_cast<float>(R);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int ty = 0; ty < 3; ++ty) {
    const int qy = py + ty - 1;
    if (qy < 0 || qy >= R) continue;
    int y0, y1;
    float l0, l1;
    bilinear_src(qy, scale, R0, y0, y1, l0, l1);
    const size_t col = static_cast<size_t>(ty * 3 + tx) * HID + c4 * 4;
    const float4 z0 = __ldg(reinterpret_cast<const float4*>(Z + ((b * 64 + y0 * 8 + c) * 9 * HID + col)));
    const float4 z1 = __ldg(reinterpret_cast<const float4*>(Z + ((b * 64 + y1 * 8 + c) * 9 * HID + col)));
    acc.x += l0 * z0.x + l1 * z1.x; acc.y += l0 * z0.y + l1 * z1.y;
    acc.z += l0 * z0.z + l1 * z1.z; acc.w += l0 * z0.w + l1 * z1.w;
  }
  reinterpret_cast<float4*>(Ry)[t] = acc;
}
// actv[b, py, px, :] = ReLU(bias + column stage) → split bf16 [B·R², HID].  A warp = the 128 channels (lane = channel
// quad) of PXW consecutive pixels: the two source columns a tap reads change at most once per 8/R0·R pixels, so their Ry
// rows stay in registers across the run and every pixel costs 6 FMAs per channel plus its 512 B of stores.
constexpr int PXW = 8;
__global__ void __launch_bounds__(256)
spade_hidden_kernel(const float* __restrict__ Ry, const float* __restrict__ bias, int R, size_t n_groups, bf16* hi,
                    bf16* lo) {
  const size_t grp = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (grp >= n_groups) return;                          // n_groups = B·R²/PXW (R ≥ 8: a group never leaves its row)
  const int c4 = threadIdx.x & 31;
  const size_t p0 = grp * PXW;
  const int lg = 31 - __clz(R);
  const int px0 = static_cast<int>(p0 & (R - 1));
  const size_t row = p0 >> lg;                          // b·R + py
  const float scale = static_cast<float>(R0) / static_cast<float>(R);
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  const float* ry = Ry + row * (3 * 8 * HID) + c4 * 4;
  float4 r0[3], r1[3];
  int cx0[3] = {-1, -1, -1}, cx1[3] = {-1, -1, -1};
  // pixel px reads the positions px − 1, px, px + 1: a sliding window, one new position per pixel (its source columns and
  // weights are the same for every lane — computing them three times per pixel was a third of this kernel's instructions)
  int wx0[3], wx1[3];
  float wl0[3], wl1[3];
  bool wok[3];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int qx = px0 + m - 1;
    wok[m + 1] = qx >= 0 && qx < R;
    bilinear_src(qx, scale, R0, wx0[m + 1], wx1[m + 1], wl0[m + 1], wl1[m + 1]);
  }
#pragma unroll
  for (int k = 0; k < PXW; ++k) {
    const int px = px0 + k;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      wok[m] = wok[m + 1]; wx0[m] = wx0[m + 1]; wx1[m] = wx1[m + 1]; wl0[m] = wl0[m + 1]; wl1[m] = wl1[m + 1];
    }
    wok[2] = px + 1 < R;
    bilinear_src(px + 1, scale, R0, wx0[2], wx1[2], wl0[2], wl1[2]);
    float4 acc = bv;
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      if (!wok[tx]) continue;
      const int x0 = wx0[tx], x1 = wx1[tx];
      const float l0 = wl0[tx], l1 = wl1[tx];
      if (x0 != cx0[tx]) { r0[tx] = __ldg(reinterpret_cast<const float4*>(ry + (tx * 8 + x0) * HID)); cx0[tx] = x0; }
      if (x1 != cx1[tx]) { r1[tx] = __ldg(reinterpret_cast<const float4*>(ry + (tx * 8 + x1) * HID)); cx1[tx] = x1; }
      acc.x += l0 * r0[tx].x + l1 * r1[tx].x; acc.y += l0 * r0[tx].y + l1 * r1[tx].y;
      acc.z += l0 * r0[tx].z + l1 * r1[tx].z; acc.w += l0 * r0[tx].w + l1 * r1[tx].w;
    }
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    store_split(hi, lo, (p0 + k) * HID + c4 * 4, acc);
  }
}

// SPADE modulation + noise + LeakyReLU(0.2) at one source pixel (layers.py:33-47, 56-62, 97-100):
//   t = ((x − mean)·rstd)·(1 + γ) + β  (+ w·noise);  t = t > 0 ? t : 0.2·t
__device__ __forceinline__ float4 spade_pixel(const float* __restrict__ x, int ldx, const float* __restrict__ gb,
                                              size_t pix, int c4, float4 mu, float4 rs, const float* noise, float nw) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + pix * ldx + c4 * 4));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gb + pix * (2 * CH) + c4 * 4));
  const float4 b = __ldg(reinterpret_cast<const float4*>(gb + pix * (2 * CH) + CH + c4 * 4));
  const float nz = noise ? nw * __ldg(noise + pix) : 0.f;
  float4 t;
  t.x = ((v.x - mu.x) * rs.x) * (1.f + g.x) + b.x + nz;
  t.y = ((v.y - mu.y) * rs.y) * (1.f + g.y) + b.y + nz;
  t.z = ((v.z - mu.z) * rs.z) * (1.f + g.z) + b.z + nz;
  t.w = ((v.w - mu.w) * rs.w) * (1.f + g.w) + b.w + nz;
  t.x = t.x > 0.f ? t.x : 0.2f * t.x; t.y = t.y > 0.f ? t.y : 0.2f * t.y;
  t.z = t.z > 0.f ? t.z : 0.2f * t.z; t.w = t.w > 0.f ? t.w : 0.2f * t.w;
  return t;
}
// out[B, up·R, up·R, CH] (split) = bilinear_up( lrelu(spade(x)) ), up ∈ {1, 2}.  With gb == nullptr the modulation is
// skipped and the kernel is a plain ×up bilinear resize of x (the residual branch, layers.py:88-91).
__global__ void spade_act_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, const float* __restrict__ gb,
                                 const float* __restrict__ noise, const float* __restrict__ noise_w, int R, int up,
                                 size_t npix_out, bf16* hi, bf16* lo) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t >= npix_out * (CH / 4)) return;
  // every extent is a power of two (CH/4 = 8, Ro = 16…256): shifts and masks instead of 64-bit divisions, which used
  // to cost more than the memory traffic of this kernel
  static_assert(CH / 4 == 8, "index math below assumes 8 channel quads per pixel");
  const int c4 = static_cast<int>(t & 7);
  const size_t p = t >> 3;
  const int Ro = R * up, lg = 31 - __clz(Ro);
  const int ox = static_cast<int>(p & (Ro - 1)), oy = static_cast<int>((p >> lg) & (Ro - 1));
  const size_t b = p >> (2 * lg);
  float4 mu = make_float4(0, 0, 0, 0), rs = make_float4(1, 1, 1, 1);
  if (gb) {
    mu = __ldg(reinterpret_cast<const float4*>(mean + b * CH + c4 * 4));
    rs = __ldg(reinterpret_cast<const float4*>(rstd + b * CH + c4 * 4));
  }
  const float nw = (noise && noise_w) ? __ldg(noise_w) : 0.f;
  const size_t base = b * R * R;
  auto fetch = [&](int y, int xx) -> float4 {
    const size_t pix = base + static_cast<size_t>(y) * R + xx;
    if (gb) return spade_pixel(x, ldx, gb, pix, c4, mu, rs, noise, nw);
    return __ldg(reinterpret_cast<const float4*>(x + pix * ldx + c4 * 4));
  };
  float4 o;
  if (up == 1) {
    o = fetch(oy, ox);
  } else {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(oy, 0.5f, R, y0, y1, ly0, ly1);
    bilinear_src(ox, 0.5f, R, x0, x1, lx0, lx1);
    const float4 v00 = fetch(y0, x0), v01 = fetch(y0, x1), v10 = fetch(y1, x0), v11 = fetch(y1, x1);
    o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
    o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
    o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
    o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  }
  store_split(hi, lo, p * CH + c4 * 4, o);
}

// ×2 variant of spade_act_kernel: a thread owns one SOURCE pixel (4 channels) and produces its 2 × 2 output pixels from
// the 3 × 3 source neighbourhood — every source value (and its SPADE modulation) is evaluated 2.25× instead of 4× per
// output and the index arithmetic is paid once per four outputs.  Same operation order per output as the generic
// kernel: output 2j / 2j+1 blends sources (j−1, j) / (j, j+1) with bilinear_src's weights; at the borders the clamped
// neighbour carries weight 0 or duplicates the centre, which is exactly what upsample_bilinear2d computes.
__global__ void __launch_bounds__(256)
spade_up2_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                 const float* __restrict__ gb, const float* __restrict__ noise, const float* __restrict__ noise_w, int R,
                 size_t npix_src, bf16* hi, bf16* lo) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t >= npix_src * (CH / 4)) return;
  const int c4 = static_cast<int>(t & 7);
  const size_t p = t >> 3;
  const int lg = 31 - __clz(R);
  const int k = static_cast<int>(p & (R - 1)), j = static_cast<int>((p >> lg) & (R - 1));
  const size_t b = p >> (2 * lg);
  float4 mu = make_float4(0, 0, 0, 0), rs = make_float4(1, 1, 1, 1);
  if (gb) {
    mu = __ldg(reinterpret_cast<const float4*>(mean + b * CH + c4 * 4));
    rs = __ldg(reinterpret_cast<const float4*>(rstd + b * CH + c4 * 4));
  }
  const float nw = (noise && noise_w) ? __ldg(noise_w) : 0.f;
  const size_t base = b * R * R;
  auto fetch = [&](int y, int xx) -> float4 {
    const size_t pix = base + static_cast<size_t>(y) * R + xx;
    if (gb) return spade_pixel(x, ldx, gb, pix, c4, mu, rs, noise, nw);
    return __ldg(reinterpret_cast<const float4*>(x + pix * ldx + c4 * 4));
  };
  const int ys[3] = {j > 0 ? j - 1 : 0, j, j < R - 1 ? j + 1 : R - 1};
  const int xs[3] = {k > 0 ? k - 1 : 0, k, k < R - 1 ? k + 1 : R - 1};
  float4 g[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) g[a][c] = fetch(ys[a], xs[c]);
  const int Ro = 2 * R;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    int i0, i1;
    float ly0, ly1;
    bilinear_src(2 * j + dy, 0.5f, R, i0, i1, ly0, ly1);
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float lx0, lx1;
      bilinear_src(2 * k + dx, 0.5f, R, i0, i1, lx0, lx1);
      const float4 v00 = g[dy][dx], v01 = g[dy][dx + 1], v10 = g[dy + 1][dx], v11 = g[dy + 1][dx + 1];
      float4 o;
      o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
      o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
      o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
      o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
      const size_t op = (b * Ro + static_cast<size_t>(2 * j + dy)) * Ro + (2 * k + dx);
      store_split(hi, lo, op * CH + c4 * 4, o);
    }
  }
}

// ToRGB tail (layers.py:126-132, 241-251): img[b, c, :, :] (+)= bilinear(rgb[b, :, :, c] → T×T); the last block
// applies tanh and optionally exports the pre-tanh sum.  rgb: [B, r, r, RGB_LD] fp32; img NCHW [B, 3, T, T].
__global__ void rgb_accumulate_kernel(const float* __restrict__ rgb, int r, int T, size_t n, int first, int last,
                                      float* acc, float* img, float* pre_tanh) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t >= n) return;
  const int lg = 31 - __clz(T);                       // T is a power of two (256)
  const int ox = static_cast<int>(t & (T - 1)), oy = static_cast<int>((t >> lg) & (T - 1));
  const unsigned plane = static_cast<unsigned>(t >> (2 * lg));   // b·3 + c
  const int c = plane % 3u;
  const size_t b = plane / 3u;
  const float* rb = rgb + b * r * r * RGB_LD + c;
  float v;
  if (r == T) {
    v = __ldg(rb + (static_cast<size_t>(oy) * r + ox) * RGB_LD);
  } else {
    const float scale = static_cast<float>(r) / static_cast<float>(T);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(oy, scale, r, y0, y1, ly0, ly1);
    bilinear_src(ox, scale, r, x0, x1, lx0, lx1);
    const float v00 = __ldg(rb + (static_cast<size_t>(y0) * r + x0) * RGB_LD);
    const float v01 = __ldg(rb + (static_cast<size_t>(y0) * r + x1) * RGB_LD);
    const float v10 = __ldg(rb + (static_cast<size_t>(y1) * r + x0) * RGB_LD);
    const float v11 = __ldg(rb + (static_cast<size_t>(y1) * r + x1) * RGB_LD);
    v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
  const float s = first ? v : acc[t] + v;      // out = zeros + rgb_0 + rgb_1 + … in block order
  if (last) {
    if (pre_tanh) pre_tanh[t] = s;
    img[t] = tanhf(s);
  } else {
    acc[t] = s;
  }
}

// ---- parameter slots ----------------------------------------------------------------------------------
// 0,1 bottleneck_emb.0.{weight,bias}; 2-5 learned_init_conv.0.{weight_orig,bias,weight_u,weight_v};
// 6-9 style_init_conv.0.{…}; block i at 10 + 26·i: cbn1.{shared.0,gamma,beta}.{weight,bias} (6), cbn2.… (6),
// conv1.{weight_orig,bias,weight_u,weight_v}, conv2.{…}, res_branch.1.{…} (12), noise1.weight, noise2.weight (2);
// to_RGB_blocks.i.conv.{weight,bias} at 140 + 2·i.
inline int blk(int i) { return 10 + 26 * i; }
inline int rgb_slot(int i) { return 10 + 26 * NBLK + 2 * i; }

// shared: [HID, 9·CH] conv layout (XLX_SPADE_ANALYTIC=0 path); shared_z: [9·HID, CH], row (tap·HID + co) — the B operand
// of Z = y·W_tᵀ (see spade_rows_kernel)
struct SpadeW { Split shared, shared_z, gb; float* gb_bias; };
struct Prep {
  Split bott, init;          // [256, 2048]; [64, 9·256] (rows 0-31 learned_init, 32-63 style_init)
  float* init_bias;          // [64]
  SpadeW sp[NBLK][2];
  Split conv1[NBLK], conv2[NBLK], rgb[NBLK];   // [64(pad), 288], [64, 288], [64, 288]
  // 1×1 shortcut convolution (layers.py:88-91) as a plain GEMM over groups of RES_PACK pixels: the weight is stored
  // block-diagonally, [RES_PACK·32, RES_PACK·32], so that [M, 32]·Wᵀ becomes [M/4, 128]·W_bdᵀ on the same memory —
  // four times fewer, four times deeper tiles (a K = 32 tile is pure pipeline latency)
  Split res[NBLK];
  float* res_bias[NBLK];     // the bias repeated RES_PACK times
  float* sig;                // 1/σ per spectrally normalised conv: 2 + 3·NBLK
  float* rgb_bias[NBLK];     // [RGB_LD]
  size_t bytes;
};
Prep prep_layout(void* base) {
  Bump b; b.base = static_cast<char*>(base);
  Prep p;
  p.bott = b.split(static_cast<size_t>(CODE) * EMB);
  p.init = b.split(static_cast<size_t>(64) * 9 * CODE);
  p.init_bias = b.f32(64);
  for (int i = 0; i < NBLK; ++i) {
    for (int j = 0; j < 2; ++j) {
      p.sp[i][j].shared = b.split(static_cast<size_t>(HID) * 9 * CH);
      p.sp[i][j].shared_z = b.split(static_cast<size_t>(9) * HID * CH);
      p.sp[i][j].gb = b.split(static_cast<size_t>(2 * CH) * 9 * HID);
      p.sp[i][j].gb_bias = b.f32(2 * CH);
    }
    p.conv1[i] = b.split(static_cast<size_t>(64) * 9 * CH);
    p.conv2[i] = b.split(static_cast<size_t>(64) * 9 * CH);
    p.res[i] = b.split(static_cast<size_t>(RES_PACK * CH) * RES_PACK * CH);
    p.res_bias[i] = b.f32(RES_PACK * CH);
    p.rgb[i] = b.split(static_cast<size_t>(64) * 9 * CH);
    p.rgb_bias[i] = b.f32(RGB_LD);
  }
  p.sig = b.f32(2 + 3 * NBLK);
  p.bytes = b.total();
  return p;
}

struct Ws {
  Split emb, code, yup[NBLK + 1], actv, a, xu, os;
  float *hy, *gb, *h1, *of[2], *rgb, *acc, *mean, *rstd, *z;
  double* stat;
  size_t bytes;
};
Ws ws_layout(int B, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  Ws w;
  const size_t n = B, top = static_cast<size_t>(256) * 256;
  w.emb = b.split(n * 64 * EMB);
  w.code = b.split(n * 64 * CODE);
  w.hy = b.f32(n * 64 * 64);
  w.z = b.f32(n * 64 * 9 * HID);          // Z_t = W_t·y at 8×8 for the nine taps (spade_rows_kernel)
  for (int i = 0; i <= NBLK; ++i) w.yup[i] = b.split(n * (static_cast<size_t>(R0 << i) * (R0 << i)) * CH);
  w.actv = b.split(n * top * HID);
  w.gb = b.f32(n * top * 2 * CH);
  w.a = b.split(n * top * CH);
  w.xu = b.split(n * top * CH);
  w.os = b.split(n * top * CH);
  w.h1 = b.f32(n * top * CH);
  w.of[0] = b.f32(n * top * CH);
  w.of[1] = b.f32(n * top / 4 * CH);
  w.rgb = b.f32(n * top * RGB_LD);
  w.acc = b.f32(n * 3 * top);
  w.mean = b.f32(n * CH);
  w.rstd = b.f32(n * CH);
  w.stat = static_cast<double*>(b.take(n * CH * 2 * sizeof(double)));
  w.bytes = b.total();
  return w;
}

// conv as implicit GEMM over an NHWC split tensor: out[B·R², N] = conv_{taps}(in[B,R,R,C]) · Wᵀ (+ epilogue)
int conv(int passes, cudaStream_t st, Split in, int B, int R, int C, int taps, Split w, int w_rows, int N,
         const GemmEpilogue& e) {
  GemmProblem p;
  p.M = B * R * R; p.N = N; p.K = taps * C; p.passes = passes;
  p.a.hi = in.hi; p.a.lo = in.lo; p.a.ld = C;
  p.conv.enabled = 1; p.conv.H = R; p.conv.W = R; p.conv.C = C; p.conv.taps = taps;
  p.b.hi = w.hi; p.b.lo = w.lo; p.b.ld = taps * C; p.b.mn_major = 0; p.b.rows = w_rows;
  p.epi = e;
  return gemm_launch(p, st);
}

int instance_stats(const float* x, int ld, int B, int HW, const Ws& w, cudaStream_t st) {
  XLX_CUDA(cudaMemsetAsync(w.stat, 0, static_cast<size_t>(B) * CH * 2 * sizeof(double), st));
  int chunks = (HW + 255) / 256;
  if (chunks > 64) chunks = 64;
  in_stats_kernel<<<dim3(chunks, B), 256, 0, st>>>(x, ld, HW, w.stat);
  XLX_TRY(krc());
  in_finish_kernel<<<(B * CH + 255) / 256, 256, 0, st>>>(w.stat, B * CH, HW, w.mean, w.rstd);
  return krc();
}

inline unsigned blocks_for(size_t n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace

extern "C" {

int64_t xlx_generator_launch_count(void) { return g_gen_launches.load(); }
int64_t xlx_generator_num_params(void) { return N_PARAMS; }
size_t xlx_generator_prep_bytes(void) { return prep_layout(nullptr).bytes; }
size_t xlx_generator_workspace_bytes(int32_t B) { return B > 0 ? ws_layout(B, nullptr).bytes : 0; }

int32_t xlx_generator_prepare(const float* const* P, void* prep, void* stream) {
  if (!P || !prep) return -24;
  XLX_TRY(ensure_device(prep));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Prep p = prep_layout(prep);
  XLX_CUDA(cudaMemsetAsync(prep, 0, p.bytes - 256, st));     // padded rows / off-group blocks stay zero
  auto prep_w = [&](const float* w, const float* scale, int Co, int Ci, int taps, int ctot, int cpg, int row0,
                    Split dst, int col_base = 0) -> int {
    const int n = Co * Ci * taps;
    prep_conv_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, scale, Co, Ci, taps, ctot, cpg, row0, taps * ctot, dst.hi, dst.lo,
                                                       col_base);
    return krc();
  };
  auto sigma = [&](int slot, int rows, int cols, float* out) -> int {
    sigma_inv_kernel<<<1, 256, 0, st>>>(P[slot], P[slot + 2], P[slot + 3], rows, cols, out);
    return krc();
  };
  XLX_TRY(prep_w(P[0], nullptr, CODE, EMB, 1, EMB, 0, 0, p.bott));
  // grouped init convs (groups = 4, layers.py:178-185): 8 output channels per group see 64 input channels
  XLX_TRY(sigma(2, CH, (CODE / 4) * 9, p.sig + 0));
  XLX_TRY(sigma(6, CH, (CODE / 4) * 9, p.sig + 1));
  XLX_TRY(prep_w(P[2], p.sig + 0, CH, CODE / 4, 9, CODE, CH / 4, 0, p.init));
  XLX_TRY(prep_w(P[6], p.sig + 1, CH, CODE / 4, 9, CODE, CH / 4, CH, p.init));
  concat_bias_kernel<<<1, 64, 0, st>>>(P[3], CH, P[7], CH, p.init_bias, 64);
  XLX_TRY(krc());
  for (int i = 0; i < NBLK; ++i) {
    const int s = blk(i);
    for (int j = 0; j < 2; ++j) {
      const int q = s + 6 * j;
      XLX_TRY(prep_w(P[q], nullptr, HID, CH, 9, CH, 0, 0, p.sp[i][j].shared));
      prep_tapmajor_kernel<<<(HID * CH * 9 + 255) / 256, 256, 0, st>>>(P[q], HID, CH, p.sp[i][j].shared_z.hi,
                                                                      p.sp[i][j].shared_z.lo);
      XLX_TRY(krc());
      XLX_TRY(prep_w(P[q + 2], nullptr, CH, HID, 9, HID, 0, 0, p.sp[i][j].gb));       // γ rows 0-31
      XLX_TRY(prep_w(P[q + 4], nullptr, CH, HID, 9, HID, 0, CH, p.sp[i][j].gb));      // β rows 32-63
      concat_bias_kernel<<<1, 64, 0, st>>>(P[q + 3], CH, P[q + 5], CH, p.sp[i][j].gb_bias, 2 * CH);
      XLX_TRY(krc());
    }
    float* sg = p.sig + 2 + 3 * i;
    XLX_TRY(sigma(s + 12, CH, CH * 9, sg + 0));
    XLX_TRY(sigma(s + 16, CH, CH * 9, sg + 1));
    XLX_TRY(sigma(s + 20, CH, CH, sg + 2));
    XLX_TRY(prep_w(P[s + 12], sg + 0, CH, CH, 9, CH, 0, 0, p.conv1[i]));
    XLX_TRY(prep_w(P[s + 16], sg + 1, CH, CH, 9, CH, 0, 0, p.conv2[i]));
    for (int g = 0; g < RES_PACK; ++g)
      XLX_TRY(prep_w(P[s + 20], sg + 2, CH, CH, 1, RES_PACK * CH, 0, g * CH, p.res[i], g * CH));
    repeat_bias_kernel<<<1, RES_PACK * CH, 0, st>>>(P[s + 21], CH, p.res_bias[i], RES_PACK * CH);
    XLX_TRY(krc());
    XLX_TRY(prep_w(P[rgb_slot(i)], nullptr, 3, CH, 9, CH, 0, 0, p.rgb[i]));
    concat_bias_kernel<<<1, 64, 0, st>>>(P[rgb_slot(i) + 1], 3, nullptr, 0, p.rgb_bias[i], RGB_LD);
    XLX_TRY(krc());
  }
  return 0;
}

int32_t xlx_generator_fwd(const float* const* P, const void* prep, int32_t B, const float* emb,
                          const float* const* noise, float* img, float* pre_tanh, float* const* block_out,
                          void* workspace, size_t workspace_bytes, int32_t passes, void* stream) {
  if (B < 1) return -21;
  if (!P || !prep || !emb || !img || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Prep p = prep_layout(const_cast<void*>(prep));
  Ws w = ws_layout(B, workspace);
  if (w.bytes > workspace_bytes) return -23;
  const int M0 = B * R0 * R0;

  // bottleneck_emb: 1×1 conv 2048 → 256 + tanh (layers.py:147-150) = GEMM over the B·64 grid cells
  XLX_TRY(split_f32(emb, w.emb, static_cast<size_t>(M0) * EMB, st));
  {
    GemmEpilogue e;
    e.bias = P[1]; e.flags = EPI_TANH; e.out_hi = w.code.hi; e.out_lo = w.code.lo; e.ld_split = CODE;
    XLX_TRY(gemm_linear(passes, st, w.emb, M0, EMB, p.bott, CODE, e));
  }
  // learned_init_conv | style_init_conv in one grouped 3×3 conv: hy[:, :32] = h, hy[:, 32:] = y (layers.py:238-239)
  {
    GemmEpilogue e;
    e.bias = p.init_bias; e.out_f32 = w.hy; e.ld_out = 64;
    XLX_TRY(conv(passes, st, w.code, B, R0, CODE, 9, p.init, 64, 64, e));
  }
  const float* y = w.hy + CH;
  // style map resized to every resolution once (SPADE resizes y to x's size, layers.py:40); the analytic hidden map
  // (default) only needs the 8×8 original as a split GEMM operand
  static const bool analytic = [] { const char* e = getenv("XLX_SPADE_ANALYTIC"); return !(e && e[0] == '0'); }();
  for (int i = 0; i <= (analytic ? 0 : NBLK); ++i) {
    const int R = R0 << i;
    const size_t npix = static_cast<size_t>(B) * R * R;
    resize_split_kernel<<<blocks_for(npix * (CH / 4)), 256, 0, st>>>(y, 64, R0, R, npix, w.yup[i].hi, w.yup[i].lo);
    XLX_TRY(krc());
  }
  // SPADE parameter maps at resolution index ri for SPADE instance (i, j): gb = [γ | β] fp32 [B,R,R,64] — or, with
  // `mod` given, the modulated + activated block input itself straight out of the γ/β convolution's epilogue (γ|β never
  // written): split bf16 into w.a (up = 1) or fp32 [B,R,R,32] into w.gb for the ×2 bilinear kernel (up = 2)
  struct Modulate { const float* x; int ldx; const float* noise; const float* noise_w; int up; };
  static const bool spade_epi = [] { const char* e = getenv("XLX_SPADE_EPILOGUE"); return !(e && e[0] == '0'); }();
  auto spade_params = [&](int i, int j, int ri, const Modulate* mod) -> int {
    const int R = R0 << ri;
    if (analytic) {
      // Z = y·W_tᵀ for all taps (one GEMM at 8×8), row stage, column stage + ReLU → actv; Ry borrows the γ/β buffer,
      // which is only written by the convolution below
      GemmEpilogue z;
      z.out_f32 = w.z; z.ld_out = 9 * HID;
      XLX_TRY(gemm_linear(passes, st, w.yup[0], M0, CH, p.sp[i][j].shared_z, 9 * HID, z));
      const size_t nr = static_cast<size_t>(B) * R * 3 * 8 * (HID / 4), ngrp = static_cast<size_t>(B) * R * R / PXW;
      spade_rows_kernel<<<blocks_for(nr), 256, 0, st>>>(w.z, R, nr, w.gb);
      XLX_TRY(krc());
      spade_hidden_kernel<<<static_cast<unsigned>((ngrp + 7) / 8), 256, 0, st>>>(w.gb, P[blk(i) + 6 * j + 1], R, ngrp,
                                                                                 w.actv.hi, w.actv.lo);
      XLX_TRY(krc());
    } else {
      GemmEpilogue e;
      e.bias = P[blk(i) + 6 * j + 1]; e.flags = EPI_RELU; e.out_hi = w.actv.hi; e.out_lo = w.actv.lo; e.ld_split = HID;
      XLX_TRY(conv(passes, st, w.yup[ri], B, R, CH, 9, p.sp[i][j].shared, HID, HID, e));
    }
    GemmEpilogue g;
    g.bias = p.sp[i][j].gb_bias;
    if (mod) {
      g.spade_x = mod->x; g.spade_ldx = mod->ldx; g.spade_mean = w.mean; g.spade_rstd = w.rstd;
      g.spade_noise = mod->noise; g.spade_noise_w = mod->noise_w;
      g.spade_hw_log2 = 2 * (31 - __builtin_clz(static_cast<unsigned>(R)));
      if (mod->up == 1) { g.out_hi = w.a.hi; g.out_lo = w.a.lo; g.ld_split = CH; }
      else { g.out_f32 = w.gb; g.ld_out = CH; }
    } else {
      g.out_f32 = w.gb; g.ld_out = 2 * CH;
    }
    return conv(passes, st, w.actv, B, R, HID, 9, p.sp[i][j].gb, 2 * CH, 2 * CH, g);
  };

  const float* x = w.hy;     // current block input, fp32 NHWC with pixel stride ldx
  int ldx = 64;
  for (int i = 0; i < NBLK; ++i) {
    const int R = R0 << i, R2 = R * 2, s = blk(i);
    const size_t npix2 = static_cast<size_t>(B) * R2 * R2;
    const float* n1 = noise ? noise[2 * i] : nullptr;
    const float* n2 = noise ? noise[2 * i + 1] : nullptr;
    // cbn1 → noise1 → LeakyReLU → ×2 bilinear (layers.py:96-101)
    XLX_TRY(instance_stats(x, ldx, B, R * R, w, st));
    const size_t npix1 = npix2 / 4;
    if (spade_epi) {
      const Modulate m1{x, ldx, n1, P[s + 24], 2};
      XLX_TRY(spade_params(i, 0, i, &m1));          // w.gb ← lrelu(spade(x)) at R×R, fp32 [B,R,R,32]
      spade_up2_kernel<<<blocks_for(npix1 * (CH / 4)), 256, 0, st>>>(w.gb, CH, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                     R, npix1, w.a.hi, w.a.lo);
    } else {
      XLX_TRY(spade_params(i, 0, i, nullptr));
      spade_up2_kernel<<<blocks_for(npix1 * (CH / 4)), 256, 0, st>>>(x, ldx, w.mean, w.rstd, w.gb, n1, P[s + 24], R, npix1,
                                                                     w.a.hi, w.a.lo);
    }
    XLX_TRY(krc());
    // residual branch input: ×2 bilinear of x (layers.py:88-91)
    spade_up2_kernel<<<blocks_for(npix1 * (CH / 4)), 256, 0, st>>>(x, ldx, nullptr, nullptr, nullptr, nullptr, nullptr, R,
                                                                   npix1, w.xu.hi, w.xu.lo);
    XLX_TRY(krc());
    // conv1 (spectral norm folded into the prepared weight)
    {
      GemmEpilogue e;
      e.bias = P[s + 13]; e.out_f32 = w.h1; e.ld_out = CH;
      XLX_TRY(conv(passes, st, w.a, B, R2, CH, 9, p.conv1[i], 64, CH, e));
    }
    // cbn2 → noise2 → LeakyReLU (layers.py:104-107)
    XLX_TRY(instance_stats(w.h1, CH, B, R2 * R2, w, st));
    if (spade_epi) {
      const Modulate m2{w.h1, CH, n2, P[s + 25], 1};
      XLX_TRY(spade_params(i, 1, i + 1, &m2));      // w.a ← lrelu(spade(h1)) as split bf16, the operand of conv2
    } else {
      XLX_TRY(spade_params(i, 1, i + 1, nullptr));
      spade_act_kernel<<<blocks_for(npix2 * (CH / 4)), 256, 0, st>>>(w.h1, CH, w.mean, w.rstd, w.gb, n2, P[s + 25], R2, 1,
                                                                     npix2, w.a.hi, w.a.lo);
      XLX_TRY(krc());
    }
    // conv2, then out = h + res_branch(x): the 1×1 conv accumulates onto conv2's output (layers.py:108-112)
    // ping-pong: blocks alternate between the two buffers; the largest (last) block uses the full-size of[0]
    float* of = ((NBLK - 1 - i) & 1) ? w.of[1] : w.of[0];
    {
      GemmEpilogue e;
      e.bias = P[s + 17]; e.out_f32 = of; e.ld_out = CH;
      XLX_TRY(conv(passes, st, w.a, B, R2, CH, 9, p.conv2[i], 64, CH, e));
      GemmEpilogue r;   // shortcut: RES_PACK pixels per GEMM row, block-diagonal weight (see Prep::res)
      const int PK = RES_PACK * CH;
      r.bias = p.res_bias[i]; r.flags = EPI_ACCUM; r.out_f32 = of; r.ld_out = PK;
      r.out_hi = w.os.hi; r.out_lo = w.os.lo; r.ld_split = PK;
      XLX_TRY(gemm_linear(passes, st, w.xu, static_cast<int>(npix2 / RES_PACK), PK, p.res[i], PK, r));
    }
    if (block_out && block_out[i])
      XLX_CUDA(cudaMemcpyAsync(block_out[i], of, npix2 * CH * 4, cudaMemcpyDeviceToDevice, st));
    // ToRGB (layers.py:126-132) and accumulation into the image (layers.py:241-251)
    {
      GemmEpilogue e;
      e.bias = p.rgb_bias[i]; e.out_f32 = w.rgb; e.ld_out = RGB_LD;
      XLX_TRY(conv(passes, st, w.os, B, R2, CH, 9, p.rgb[i], 64, RGB_LD, e));
      const size_t n = static_cast<size_t>(B) * 3 * 256 * 256;
      rgb_accumulate_kernel<<<blocks_for(n), 256, 0, st>>>(w.rgb, R2, 256, n, i == 0, i == NBLK - 1, w.acc, img, pre_tanh);
      XLX_TRY(krc());
    }
    x = of; ldx = CH;
  }
  return 0;
}

}  // extern "C"
