// tcgen05 / TMA / TMEM GEMM for sm_100a — see gemm_sm100.cuh for the contract.
#include "gemm_sm100.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "pdl.cuh"
#include "xlx_ptx.cuh"

namespace xlx {

namespace {

constexpr int BM = 128;            // UMMA M (one TMEM lane per output row)
#ifndef XLX_EPI_WARPS
#define XLX_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = XLX_EPI_WARPS;  // EPI_WARPS/4 warps per TMEM lane quadrant, alternating 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;  // warp 0: TMA, warp 1: MMA, then the epilogue warps
constexpr int EPI_COLS = 16;       // accumulator columns per epilogue step (one tcgen05.ld.32x32b.x16)
// Per-warp transpose buffer: 32 rows × 16 fp32, no padding; the 16-byte chunk index is XOR-ed with (row >> 1) & 3,
// which makes both the row-wise writes and the 4-lanes-per-row reads bank-conflict free.
constexpr int STAGE_BYTES = EPI_WARPS * 32 * EPI_COLS * 4;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA on sm_100
constexpr size_t kSplitkCounters = 1024;   // words at the end of a split-K workspace holding the per-tile counters

struct KParams {
  int M, N, K;
  int BN;
  int nparts;  // 1 (hi only) or 2 (hi + lo)
  int a_mn, b_mn;
  int num_stages;
  int tiles_m, tiles_n;
  uint32_t stage_bytes, a_part_bytes, b_part_bytes;
  uint32_t tmem_cols;
  int splits;          // split-K factor (1 = none); work item = (tile, split)
  int kb_per_split;    // k-blocks per split (last split may be shorter)
  float* part;         // [splits][M][N] fp32 partial sums when splits > 1
  int* tile_counter;   // split-K: arrivals per output tile; the last arrival reduces the tile (nullptr: reduce kernel)
  int conv, conv_H, conv_W, conv_taps, conv_kb_per_tap;   // implicit-GEMM convolution (see ConvGeometry)
  // conv == 2 ("row halo", 3×3, W % 128 == 0): a k-block is (filter row dy, 32-channel block); its A stage is ONE haloed
  // image row segment of 130 pixels, and the three dx taps are three UMMA descriptor views of it shifted by one pixel
  // (64 bytes) each — every activation byte crosses L2→SM 3 times instead of 9.
  // conv == 3 ("row group", halo geometry): a work item is G vertically adjacent 128-pixel tiles with G accumulators in
  // TMEM.  Per 32-channel block the producer streams the G + 2 haloed image rows ONCE (each feeds up to three
  // (tile, filter-row) pairs) and the three filter-row weight groups ONCE (each serves G tiles) through a second,
  // longer-lived ring — 2.8× fewer bytes from L2 per pixel than conv == 2 and 54 instead of 18 MMAs per A stage.
  int conv_G;            // tiles per item (4 for BN ≤ 64)
  uint32_t b_group_bytes; // one weight group: 3 dx taps × BN rows × 32 channels, per part
  uint32_t b_ring_off;   // byte offset of the weight ring behind the A ring
  int b_slots;           // weight-ring slots
  uint32_t stage_tx;   // bytes one stage receives by TMA (= stage_bytes except in halo mode, whose A box is 130 rows)
  int conv_tapbox;     // halo mode: the B maps are rank-3 (channel, row, dx) and one box fetches all three dx taps
  int cluster;         // 1, or 2 = CTA pairs on adjacent m-tiles sharing the B tile by TMA multicast
  int tiles_per_split; // tiles (cluster = 1) or tile pairs (cluster = 2) per split
  int num_items;       // work items of the launch: tiles_per_split · splits
  int tma_out;         // 1: the epilogue is a plain fp32 store and goes out through TMA (mapOut) instead of st.global
  int debug;           // XLX_GEMM_DEBUG bit 0: skip the epilogue's global traffic (mainloop-only timing experiments)
  GemmEpilogue epi;
};

__device__ __forceinline__ float gelu_erf_grad(float x) {
  // d/dx [x·Φ(x)] = Φ(x) + x·φ(x)
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GeLU (erf form) and its derivative from one exponential:
//   Φ(x) = ½·erfc(−x/√2),  erfc(z) = e^{−z²}·(a₁t + … + a₅t⁵), t = 1/(1 + p·z), z ≥ 0   (Abramowitz–Stegun 7.1.26,
//   |error| ≤ 1.5e-7 on erf, i.e. fp32-epsilon class on Φ), and e^{−z²} = e^{−x²/2} is exactly what φ(x) needs.
//   gelu(x) = x·Φ(x),  gelu'(x) = Φ(x) + x·φ(x).   ≈ 20 instructions instead of erff + expf.
// The two special-function units are used raw (ex2.approx.ftz / rcp.approx.ftz): __expf and __fdividef wrap the same
// instructions in range fix-ups (2 compares + 3 predicated multiplies per element) that these arguments never need —
// the reciprocal's argument is ≥ 1, and an exponent that underflows should flush to zero anyway.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_and_grad(float x, float& y, float& dy) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  const float e = ex2_approx(x * x * -0.72134752044448170368f);          // e^{−x²/2} = 2^{−x²/2·log2(e)}
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);        // the ½ of ½·erfc folded into the coefficients
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float half_erfc = poly * t * e;                     // ½·erfc(|x|/√2) = Φ(−|x|)
  const float cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  y = x * cdf;
  dy = fmaf(x * e, 0.39894228040143267794f, cdf);
}

// Global inputs of the epilogue for 4 consecutive columns of one row, fetched ahead of the arithmetic so that the loads
// of a whole 32×16 chunk are in flight together.
// `a` holds the first fp32 input the epilogue needs (u_in, else addend, else the old output for EPI_ACCUM); the rare
// epilogues needing more than one of them load the others late.  `h`/`l`: split-bf16 addend.
struct EpiIn { float4 a; uint2 h, l; };
__device__ __forceinline__ void epilogue_fetch(const KParams& P, int row, int n, EpiIn& in) {
  const GemmEpilogue& E = P.epi;
  if ((E.flags & EPI_MUL) && E.u_in16) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(E.u_in16 + static_cast<size_t>(row) * E.ld_u + n));
    in.a = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
  } else if (E.flags & (EPI_MUL | EPI_GELU_GRAD))
    in.a = __ldg(reinterpret_cast<const float4*>(E.u_in + static_cast<size_t>(row) * E.ld_u + n));
  else if (E.addend)
    in.a = __ldg(reinterpret_cast<const float4*>(E.addend + static_cast<size_t>(row) * E.ld_addend + n));
  else if (E.out_f32 && (E.flags & EPI_ACCUM))
    in.a = *reinterpret_cast<const float4*>(E.out_f32 + static_cast<size_t>(row) * E.ld_out + n);
  if (E.addend_hi) {
    const size_t idx = static_cast<size_t>(row) * E.ld_addend + n;
    in.h = __ldg(reinterpret_cast<const uint2*>(E.addend_hi + idx));
    in.l = __ldg(reinterpret_cast<const uint2*>(E.addend_lo + idx));
  }
}

// Epilogue on 4 consecutive columns of one output row.  Called in the "coalesced domain": 4 adjacent lanes hold 16
// consecutive columns of the same row, so every global access below covers whole 32-byte sectors (64-byte fp32 /
// 32-byte bf16 segments per row).
template <bool ACT>
__device__ __forceinline__ float4 epilogue_vec4(const KParams& P, int row, int n, float4 acc, const EpiIn& in) {
  const GemmEpilogue& E = P.epi;
  float v[4] = {acc.x * E.alpha, acc.y * E.alpha, acc.z * E.alpha, acc.w * E.alpha};
  if (E.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(E.bias + n));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ACT && (E.flags & EPI_GELU)) {
    float dg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = v[j];
      gelu_and_grad(x, v[j], dg[j]);
      if (!(E.flags & EPI_SAVE_DGELU)) dg[j] = x;
    }
    if (E.out_u)
      *reinterpret_cast<float4*>(E.out_u + static_cast<size_t>(row) * E.ld_u + n) = make_float4(dg[0], dg[1], dg[2], dg[3]);
    if (E.out_u16)
      *reinterpret_cast<uint2*>(E.out_u16 + static_cast<size_t>(row) * E.ld_u + n) =
          make_uint2(pack_bf16x2(__float2bfloat16_rn(dg[0]), __float2bfloat16_rn(dg[1])),
                     pack_bf16x2(__float2bfloat16_rn(dg[2]), __float2bfloat16_rn(dg[3])));
  } else if (E.out_u) {
    *reinterpret_cast<float4*>(E.out_u + static_cast<size_t>(row) * E.ld_u + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (ACT && (E.flags & EPI_TANH)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = tanhf(v[j]);
  }
  if (ACT && (E.flags & EPI_RELU)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (ACT && E.drop.threshold) {
    const float4 m = drop_hidden4(E.drop, static_cast<size_t>(row), P.N, n);
    v[0] *= m.x; v[1] *= m.y; v[2] *= m.z; v[3] *= m.w;
  }
  const bool has_u = (E.flags & (EPI_MUL | EPI_GELU_GRAD)) != 0;
  if (ACT && (E.flags & EPI_GELU_GRAD)) {
    v[0] *= gelu_erf_grad(in.a.x); v[1] *= gelu_erf_grad(in.a.y); v[2] *= gelu_erf_grad(in.a.z); v[3] *= gelu_erf_grad(in.a.w);
  }
  if (E.flags & EPI_MUL) { v[0] *= in.a.x; v[1] *= in.a.y; v[2] *= in.a.z; v[3] *= in.a.w; }
  if (E.addend) {
    const float4 add = has_u ? __ldg(reinterpret_cast<const float4*>(E.addend + static_cast<size_t>(row) * E.ld_addend + n))
                             : in.a;
    v[0] += add.x; v[1] += add.y; v[2] += add.z; v[3] += add.w;
  }
  if (E.addend_hi) {
    v[0] += __uint_as_float(in.h.x << 16) + __uint_as_float(in.l.x << 16);
    v[1] += __uint_as_float(in.h.x & 0xffff0000u) + __uint_as_float(in.l.x & 0xffff0000u);
    v[2] += __uint_as_float(in.h.y << 16) + __uint_as_float(in.l.y << 16);
    v[3] += __uint_as_float(in.h.y & 0xffff0000u) + __uint_as_float(in.l.y & 0xffff0000u);
  }
  if (E.out_f32) {
    if (E.flags & EPI_ACCUM) {
      const float4 old = (has_u || E.addend)
                             ? *reinterpret_cast<const float4*>(E.out_f32 + static_cast<size_t>(row) * E.ld_out + n)
                             : in.a;
      v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
    }
    *reinterpret_cast<float4*>(E.out_f32 + static_cast<size_t>(row) * E.ld_out + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (E.out_hi) {
    const size_t idx = static_cast<size_t>(row) * E.ld_split + n;
    uint2 hw, lw;
    split_bf16x2(v[0], v[1], hw.x, lw.x);
    split_bf16x2(v[2], v[3], hw.y, lw.y);
    *reinterpret_cast<uint2*>(E.out_hi + idx) = hw;
    if (E.out_lo) *reinterpret_cast<uint2*>(E.out_lo + idx) = lw;
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

// ---- specialised epilogues (EPI 6-9) -----------------------------------------------------------------------------------
// The generic epilogue decides everything at run time (flags, optional pointers): 268 SASS instructions per 32-lane × 4-
// column step for the FFN-1 forward, which made that GEMM EPILOGUE-bound (ncu: 83 K warp instructions per 128 × 256 tile,
// issue-limited at 24 µs against a 13.6 µs main loop; profiles/r02_ffn1_epilogue_ncu.json).  The four configurations that
// carry most of the training step's epilogue work get compile-time bodies: all four rows of a chunk step in flight
// together, the bias fetched once per chunk, bf16x2 conversions.
//   6 GELU_FWD   v = acc + bias;  out_u = gelu'(v) (fp32, training only);  split(gelu(v)) → out_hi/out_lo   FFN-1 forward
//   7 SPLIT      v = acc (+ bias);  split(v) → out_hi/out_lo                       Q/K/V projections, most dgrads
//   8 RESID_F32  v = acc (+ bias) + (addend_hi + addend_lo) → out_f32    attention-output / FFN-2 forward (no dropout),
//                                                                        dgrads that add the residual path's gradient
//   9 MUL_SPLIT  v = acc · u_in (fp32);  split(v) → out_hi/out_lo                       FFN-2 dgrad × saved gelu'
//  10 RESID_DROP v = dropout(acc + bias) + (addend_hi + addend_lo) → out_f32          the same two in training mode
// Preconditions (checked by the host): alpha = 1, no dropout, no column sums, single split, K-major operands, 3 passes.
template <int EPI>
__device__ __forceinline__ void lean_chunk(const KParams& P, const float* stage, int row0, int n, int sub, int cq) {
  const GemmEpilogue& E = P.epi;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI != 9 && E.bias) bias = __ldg(reinterpret_cast<const float4*>(E.bias + n));
  float4 acc[4];
  bool ok[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + sub;
    ok[it] = row0 + rr < P.M;
    acc[it] = *reinterpret_cast<const float4*>(stage + rr * EPI_COLS + ((cq ^ ((rr >> 1) & 3)) << 2));
  }
  if (EPI == 8 || EPI == 10) {   // residual as split bf16: all eight loads first
    uint2 h[4], l[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (!ok[it]) continue;
      const size_t idx = static_cast<size_t>(row0 + it * 8 + sub) * E.ld_addend + n;
      h[it] = __ldg(reinterpret_cast<const uint2*>(E.addend_hi + idx));
      l[it] = __ldg(reinterpret_cast<const uint2*>(E.addend_lo + idx));
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (!ok[it]) continue;
      float4 v = acc[it];
      v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
      if (EPI == 10) {     // dropout(dense(x)) + residual (HF:282-287,344-349): the mask is a function of (site, row, column)
        const float4 m = drop_hidden4(E.drop, static_cast<size_t>(row0 + it * 8 + sub), P.N, n);
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
      }
      v.x += __uint_as_float(h[it].x << 16) + __uint_as_float(l[it].x << 16);
      v.y += __uint_as_float(h[it].x & 0xffff0000u) + __uint_as_float(l[it].x & 0xffff0000u);
      v.z += __uint_as_float(h[it].y << 16) + __uint_as_float(l[it].y << 16);
      v.w += __uint_as_float(h[it].y & 0xffff0000u) + __uint_as_float(l[it].y & 0xffff0000u);
      *reinterpret_cast<float4*>(E.out_f32 + static_cast<size_t>(row0 + it * 8 + sub) * E.ld_out + n) = v;
    }
    return;
  }
  float4 u[4];
  if (EPI == 9) {
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (ok[it]) u[it] = __ldg(reinterpret_cast<const float4*>(E.u_in + static_cast<size_t>(row0 + it * 8 + sub) * E.ld_u + n));
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (!ok[it]) continue;
    const int row = row0 + it * 8 + sub;
    float4 v = acc[it];
    if (EPI == 9) {
      v.x *= u[it].x; v.y *= u[it].y; v.z *= u[it].z; v.w *= u[it].w;
    } else {
      v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
    }
    if (EPI == 6) {
      float4 dg;
      gelu_and_grad(v.x, v.x, dg.x); gelu_and_grad(v.y, v.y, dg.y);
      gelu_and_grad(v.z, v.z, dg.z); gelu_and_grad(v.w, v.w, dg.w);
      if (E.out_u) *reinterpret_cast<float4*>(E.out_u + static_cast<size_t>(row) * E.ld_u + n) = dg;
    }
    uint2 hw, lw;
    split_bf16x2(v.x, v.y, hw.x, lw.x);
    split_bf16x2(v.z, v.w, hw.y, lw.y);
    const size_t idx = static_cast<size_t>(row) * E.ld_split + n;
    *reinterpret_cast<uint2*>(E.out_hi + idx) = hw;
    *reinterpret_cast<uint2*>(E.out_lo + idx) = lw;
  }
}

// A_MN / B_MN: operand stored MN-major; NPARTS: 1 = hi only (1 pass), 2 = hi + lo (3 passes).  Compile-time so that
// the single MMA-issuing thread's loop is a handful of integer adds per tcgen05.mma (it is the critical path).
// ACT: the epilogue may contain an activation (GeLU / tanh / ReLU / gelu-grad); kept out of the other instantiations so
// that their code stays small — the epilogue warps share the instruction cache with the MMA-issuing warp.
// EPI: 0 = load-bound epilogues (the common case), 1 = ACT, 2 = row statistics (soft-max / arg-max of the sampler's
// logits GEMM, K-major operands only), 3 = split-K with the reduction folded into the last-arriving CTA (weight-gradient
// layout only; XLX_GEMM_SPLITK_FOLD=1), 4 = SPADE modulation (the generator's γ/β convolution, K-major, N = BN = 64).  The rare modes are separate instantiations so that their state (16 extra live
// registers for the row statistics, the extra barriers of the fold) cannot push the hot variant into spilling.
template <int BK, int A_MN, int B_MN, int NPARTS, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
            const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
            const __grid_constant__ CUtensorMap mapOut, const KParams P) {
  constexpr bool ACT = (EPI == 1);
  constexpr bool ROWSTAT = (EPI == 2);
  constexpr bool FOLD = (EPI == 3);
  constexpr bool SPADE = (EPI == 4);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int splitk_last;     // split-K: did this CTA store the last partial of the tile?
  __shared__ __align__(8) uint64_t bar_bfull[4];    // conv == 3: weight-group ring
  __shared__ __align__(8) uint64_t bar_bempty[4];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms need 1024B alignment

  const int nkb = (P.conv == 2) ? 3 * P.conv_kb_per_tap : (P.K + BK - 1) / BK;
  // Work decomposition.  cluster = 1: item → (tile, split).  cluster = 2: the two CTAs of a cluster take the m-tiles
  // 2p and 2p + 1 of the same n-tile, so that each loads half of the shared B tile and multicasts it to both.
  const int crank = (P.cluster == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int item0 = (P.cluster == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int istride = (P.cluster == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_items = P.num_items;
  auto decode = [&](int item, int& m0, int& n0, int& split) {
    const int t = item % P.tiles_per_split;
    split = item / P.tiles_per_split;
    const int mt = (P.cluster == 2) ? 2 * (t / P.tiles_n) + crank : t / P.tiles_n;
    m0 = mt * BM;
    n0 = (t % P.tiles_n) * P.BN;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapAhi);
    tma_prefetch_desc(&mapBhi);
    if (NPARTS == 2) {
      tma_prefetch_desc(&mapAlo);
      tma_prefetch_desc(&mapBlo);
    }
    for (int s = 0; s < P.num_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), P.cluster);     // one tcgen05.commit per CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(smem_u32(&bar_bfull[b]), 1);
      mbar_init(smem_u32(&bar_bempty[b]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (P.cluster == 2) cluster_sync_all();     // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // Everything above (barriers, TMEM, descriptor prefetch) may overlap the previous kernel's tail (see launch_pdl).
  pdl_trigger();
  pdl_wait();

  // conv == 3: item → (image, first row of the G-row group, first pixel of the 128-pixel column block)
  auto decode_group = [&](int item, int& img, int& y0, int& x0) {
    const int xt = P.conv_W / BM, yg = P.conv_H / P.conv_G;
    x0 = (item % xt) * BM;
    y0 = ((item / xt) % yg) * P.conv_G;
    img = item / (xt * yg);
  };
  if (warp == 0 && !A_MN && !B_MN && P.conv == 3) {
    // ===================== TMA producer, row-group convolution =====================
    if (lane == 0) {
      int s = 0, bs = 0;
      uint32_t ph = 0, bph = 0;
      const uint32_t a_box = (BM + 2) * BK * 2;
      for (int item = item0; item < num_items; item += istride) {
        int img, y0, x0;
        decode_group(item, img, y0, x0);
        for (int cb = 0; cb < P.conv_kb_per_tap; ++cb) {
          for (int r = 0; r < P.conv_G + 2; ++r) {
            if (r < 3) {     // weight group (cb, filter row r): needed from image row r on
              mbar_wait(smem_u32(&bar_bempty[bs]), bph ^ 1);
              const uint32_t bfull = smem_u32(&bar_bfull[bs]);
              mbar_arrive_expect_tx(bfull, NPARTS * P.b_group_bytes);
              const uint32_t dB = smem_base + P.b_ring_off + bs * NPARTS * P.b_group_bytes;
#pragma unroll
              for (int part = 0; part < NPARTS; ++part)
                tma_load_3d(dB + part * P.b_group_bytes, part ? &mapBlo : &mapBhi, bfull,
                            (r * 3 * P.conv_kb_per_tap + cb) * BK, 0, 0);
              if (++bs == P.b_slots) { bs = 0; bph ^= 1; }
            }
            mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
            const uint32_t full = smem_u32(&bar_full[s]);
            mbar_arrive_expect_tx(full, NPARTS * a_box);
            const uint32_t sA = smem_base + s * P.stage_bytes;
#pragma unroll
            for (int part = 0; part < NPARTS; ++part)
              tma_load_4d(sA + part * P.a_part_bytes, part ? &mapAlo : &mapAhi, full, cb * BK, x0 - 1, y0 + r - 1, img);
            if (++s == P.num_stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = item0; item < num_items; item += istride) {
        int m0, n0, split;
        decode(item, m0, n0, split);
        const int kb0 = split * P.kb_per_split, kb1 = min(nkb, kb0 + P.kb_per_split);
        // conv mode: tile → (image, y, x) of its first pixel
        int img = 0, y0 = 0, x0 = 0;
        if (P.conv) {
          const int hw = P.conv_H * P.conv_W;
          img = m0 / hw;
          const int rem = m0 % hw;
          y0 = rem / P.conv_W; x0 = rem % P.conv_W;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          mbar_arrive_expect_tx(full, P.stage_tx);
          const uint32_t sA = smem_base + s * P.stage_bytes;
          const uint32_t sB = sA + NPARTS * P.a_part_bytes;
          const int k0 = kb * BK;
#pragma unroll
          for (int part = 0; part < NPARTS; ++part) {
            const CUtensorMap* ma = part ? &mapAlo : &mapAhi;
            const CUtensorMap* mb = part ? &mapBlo : &mapBhi;
            const uint32_t dA = sA + part * P.a_part_bytes;
            const uint32_t dB = sB + part * P.b_part_bytes;
            if (!A_MN && !B_MN && P.conv == 2) {
              // k-block → (filter row, channel block): one haloed row segment [x0 − 1, x0 + 129) and the 3 weight taps
              const int dyi = kb / P.conv_kb_per_tap, cb = kb % P.conv_kb_per_tap;
              tma_load_4d(dA, ma, full, cb * BK, x0 - 1, y0 + dyi - 1, img);
              if (P.conv_tapbox) {   // one request for the three dx taps (the TMA unit is request-rate bound here)
                tma_load_3d(dB, mb, full, (dyi * 3 * P.conv_kb_per_tap + cb) * BK, n0, 0);
              } else {
#pragma unroll
                for (int dxi = 0; dxi < 3; ++dxi)
                  tma_load_2d(dB + dxi * (P.BN * BK * 2), mb, full, ((dyi * 3 + dxi) * P.conv_kb_per_tap + cb) * BK, n0);
              }
              continue;
            } else if (!A_MN && P.conv) {
              // k-block → (filter tap, channel offset); the box is the activation tensor shifted by the tap
              const int tap = kb / P.conv_kb_per_tap, c0 = (kb % P.conv_kb_per_tap) * BK;
              const int dy = P.conv_taps == 9 ? tap / 3 - 1 : 0, dx = P.conv_taps == 9 ? tap % 3 - 1 : 0;
              tma_load_4d(dA, ma, full, c0, x0 + dx, y0 + dy, img);
            } else if (!A_MN) {
              tma_load_2d(dA, ma, full, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(dA + j * (BK * 128), ma, full, m0 + 64 * j, k0);
            }
            if (P.cluster == 2) {
              // this CTA fetches its half of the B tile and multicasts it into both CTAs' stage s
              if (!B_MN) {
                const int half_rows = P.BN / 2;
                tma_load_2d_mcast(dB + crank * half_rows * (BK * 2), mb, full, k0, n0 + crank * half_rows, 3);
              } else {
                const int nbox = P.BN / 128;
                for (int j = crank * nbox; j < (crank + 1) * nbox; ++j)
                  tma_load_2d_mcast(dB + j * (BK * 128), mb, full, n0 + 64 * j, k0, 3);
              }
            } else if (!B_MN) {
              tma_load_2d(dB, mb, full, k0, n0);
            } else {
              for (int j = 0; j < P.BN / 64; ++j) tma_load_2d(dB + j * (BK * 128), mb, full, n0 + 64 * j, k0);
            }
          }
          if (++s == P.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (whole warp runs the loop, one elected lane issues) =====================
    const uint32_t idesc = umma_idesc_bf16(BM, P.BN, A_MN, B_MN);
    // Shared-memory descriptors: the high word is a per-operand constant, the low word is
    // (LBO >> 4) << 16 | (address >> 4); stepping along K only adds to the address field.
    //   K-major: rows of BK·2 bytes (64B swizzle for BK = 32), 8-row groups SBO apart, K step = 32 bytes.
    //   MN-major: 64-element (128B) MN atoms = one TMA box of BK rows each (LBO apart); 8-row K groups 1024B apart.
    constexpr uint32_t kSwzK = (BK == 64) ? UMMA_SWZ_128B : UMMA_SWZ_64B;
    constexpr uint32_t kHiK = umma_desc_hi(8 * BK * 2, kSwzK), kHiMN = umma_desc_hi(1024, UMMA_SWZ_128B);
    constexpr uint32_t kLoK = umma_desc_lo_lbo(16), kLoMN = umma_desc_lo_lbo(BK * 128);
    constexpr uint32_t kStepK = 32 >> 4, kStepMN = (16 * 128) >> 4;
    constexpr uint32_t hiA = A_MN ? kHiMN : kHiK, hiB = B_MN ? kHiMN : kHiK;
    constexpr uint32_t loA = A_MN ? kLoMN : kLoK, loB = B_MN ? kLoMN : kLoK;
    constexpr uint32_t stepA = A_MN ? kStepMN : kStepK, stepB = B_MN ? kStepMN : kStepK;
    const uint32_t a_part = P.a_part_bytes >> 4, b_part = P.b_part_bytes >> 4;
    int s = 0;
    uint32_t ph = 0;
    int local = 0;
    if (!A_MN && !B_MN && P.conv == 3) {
      // row-group convolution: A stage = (channel block cb, image row r of the group's G + 2 haloed rows); it feeds the
      // tiles g = r − dy for the filter rows dy whose tile exists, each with the three dx taps as shifted views
      const int G = P.conv_G, kbpt = P.conv_kb_per_tap;
      const uint32_t b_tap = (P.BN * BK * 2) >> 4;
      uint32_t q = 0;                       // running index of the weight group (cb, dy = 0) of the current item
      for (int item = item0; item < num_items; item += istride, ++local) {
        const int buf = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(smem_u32(&bar_tmem_empty[buf]), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_item = tmem_base + buf * (G * P.BN);
        for (int cb = 0; cb < kbpt; ++cb, q += 3) {
          for (int r = 0; r < G + 2; ++r) {
            if (r < 3) {                    // first use of weight group (cb, dy = r)
              const uint32_t qi = q + r;
              mbar_wait(smem_u32(&bar_bfull[qi % P.b_slots]), (qi / P.b_slots) & 1);
            }
            mbar_wait(smem_u32(&bar_full[s]), ph);
            tc_fence_after();
            const uint32_t sA = (smem_base + s * P.stage_bytes) & 0x3FFFFu;
            const uint32_t a0 = loA | (sA >> 4);
            if (elect_one()) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const int g = r - dy;
                if (g < 0 || g >= G) continue;
                const uint32_t qi = q + dy;
                const uint32_t sB = (smem_base + P.b_ring_off + (qi % P.b_slots) * NPARTS * P.b_group_bytes) & 0x3FFFFu;
                const uint32_t b0 = loB | (sB >> 4);
                const uint32_t tmem_d = tmem_item + g * P.BN;
#pragma unroll
                for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
                  for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint32_t acc = (cb > 0 || dy > 0 || dxi > 0 || kk > 0) ? 1u : 0u;
                    const uint32_t ah = a0 + dxi * ((BK * 2) >> 4) + kk * stepA, bh = b0 + dxi * b_tap + kk * stepB;
                    if (NPARTS == 2) {
                      umma_bf16(tmem_d, umma_desc(hiA, ah + a_part), umma_desc(hiB, bh), idesc, acc);
                      umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh + (P.b_group_bytes >> 4)), idesc, 1u);
                      umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, 1u);
                    } else {
                      umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, acc);
                    }
                  }
                }
                if (g == G - 1) umma_commit(smem_u32(&bar_bempty[qi % P.b_slots]));   // last tile served by this group
              }
              umma_commit(smem_u32(&bar_empty[s]));
              if (cb == kbpt - 1 && r == G + 1) umma_commit(smem_u32(&bar_tmem_full[buf]));
            }
            __syncwarp();
            if (++s == P.num_stages) { s = 0; ph ^= 1; }
          }
        }
      }
    } else
    for (int item = item0; item < num_items; item += istride, ++local) {
      const int split = item / P.tiles_per_split;
      const int kb0 = split * P.kb_per_split, kb1 = min(nkb, kb0 + P.kb_per_split);
      const int buf = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      mbar_wait(smem_u32(&bar_tmem_empty[buf]), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * P.BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        // CTA-local byte offset of the stage: in a cluster launch shared-window addresses carry the CTA rank in
        // their upper bits, which must not leak into the descriptor's LBO field
        const uint32_t sA = (smem_base + s * P.stage_bytes) & 0x3FFFFu;
        const uint32_t a0 = loA | (sA >> 4);
        const uint32_t b0 = loB | ((sA + NPARTS * P.a_part_bytes) >> 4);
        if (!A_MN && !B_MN && P.conv == 2) {
          if (elect_one()) {
            const uint32_t b_tap = (P.BN * BK * 2) >> 4;
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
              // View of the haloed row shifted by dxi pixels = dxi · 64 bytes added to the descriptor's start address.
              // Measured on B200: the 64-byte swizzle is applied to absolute shared-memory address bits, so a start
              // that is not aligned to the 512-byte swizzle period reads the TMA-written tile correctly with the
              // base-offset field (bits 49-51) left at 0; filling it with (addr >> 7) & 7 gives wrong results.
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) {
                const uint32_t acc = (kb > kb0 || dxi > 0 || kk > 0) ? 1u : 0u;
                const uint32_t ah = a0 + dxi * ((BK * 2) >> 4) + kk * stepA, bh = b0 + dxi * b_tap + kk * stepB;
                if (NPARTS == 2) {
                  umma_bf16(tmem_d, umma_desc(hiA, ah + a_part), umma_desc(hiB, bh), idesc, acc);
                  umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh + b_part), idesc, 1u);
                  umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, 1u);
                } else {
                  umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, acc);
                }
              }
            }
            umma_commit(smem_u32(&bar_empty[s]));
            if (kb == kb1 - 1) umma_commit(smem_u32(&bar_tmem_full[buf]));
          }
        } else if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t acc = (kb > kb0 || kk > 0) ? 1u : 0u;
            const uint32_t ah = a0 + kk * stepA, bh = b0 + kk * stepB;
            if (NPARTS == 2) {
              // small cross terms first, leading term last
              umma_bf16(tmem_d, umma_desc(hiA, ah + a_part), umma_desc(hiB, bh), idesc, acc);
              umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh + b_part), idesc, 1u);
              umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, 1u);
            } else {
              umma_bf16(tmem_d, umma_desc(hiA, ah), umma_desc(hiB, bh), idesc, acc);
            }
          }
          // smem slot reusable once these MMAs retire (in a cluster: once BOTH CTAs' MMAs on it retired)
          if (P.cluster == 2) umma_commit_mcast(smem_u32(&bar_empty[s]), 3);
          else umma_commit(smem_u32(&bar_empty[s]));
          if (kb == kb1 - 1) umma_commit(smem_u32(&bar_tmem_full[buf]));
        }
        __syncwarp();
        if (++s == P.num_stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // TMEM → registers (thread = row) → shared-memory transpose → coalesced global traffic (4 lanes = 64 B of one
    // row, 8 rows per instruction).  The EPI_WARPS/4 warps sharing TMEM lane quadrant q = warp % 4 alternate
    // EPI_COLS-column chunks.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int e = warp - 2;                 // epilogue warp index 0 … EPI_WARPS-1
    const int half = e >> 2;                // which of the EPI_WARPS/4 warps of the quadrant
    float* stage = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + P.num_stages * P.stage_bytes +
                                            (P.conv == 3 ? P.b_slots * P.nparts * P.b_group_bytes : 0u)) +
                   e * 32 * EPI_COLS;
    const int sub = lane >> 2, cq = lane & 3;
    int local = 0;
    for (int item = item0; item < num_items; item += istride, ++local) {
      int m0, n0, split;
      // conv == 3: the item holds ngrp accumulators, one per image row of the group (tiles W pixels apart)
      const int ngrp = (P.conv == 3) ? P.conv_G : 1;
      if (P.conv == 3) {
        int img, y0, x0;
        decode_group(item, img, y0, x0);
        m0 = (img * P.conv_H + y0) * P.conv_W + x0; n0 = 0; split = 0;
      } else {
        decode(item, m0, n0, split);
      }
      const int m0_item = m0;
      const int buf = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      mbar_wait(smem_u32(&bar_tmem_full[buf]), acc_ph);
      tc_fence_after();
      const int nchunks = P.BN / EPI_COLS;
      // last chunk this warp will read (chunks beyond N are skipped)
      int last_c = -1;
      for (int c = half; c < nchunks; c += EPI_WARPS / 4)
        if (n0 + c * EPI_COLS < P.N) last_c = c;
      if (!SPADE && last_c < 0) {   // nothing to read from this accumulator: hand it back right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
      }
      // row-statistics epilogue state (thread = row domain): running max, Σ exp(x − max), first index of the max
      float rs_m = -INFINITY, rs_s = 0.f;
      int rs_arg = 0x7fffffff;
      for (int grp = 0; grp < ngrp; ++grp) {
      m0 = m0_item + grp * P.conv_W;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (ngrp * P.BN) + grp * P.BN;
      if (SPADE) {
        // thread = pixel row; the warps with half < 2 take channels half·16 … +15: their γ (columns c0 …) and β
        // (columns 32 + c0 …) accumulators, the other two warps of the quadrant only hand the accumulator back
        uint32_t rg[EPI_COLS], rb[EPI_COLS];
        if (half < 2) {
          tmem_ld_32x16(taddr + half * EPI_COLS, rg);
          tmem_ld_32x16(taddr + 32 + half * EPI_COLS, rb);
          tmem_ld_wait();
        }
        if (grp == ngrp - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
        }
        if (half < 2) {
          const GemmEpilogue& E = P.epi;
          const int row = m0 + q * 32 + lane, c0 = half * EPI_COLS;
          if (row < P.M) {
            const size_t b = static_cast<size_t>(row) >> E.spade_hw_log2;
            const float* xr = E.spade_x + static_cast<size_t>(row) * E.spade_ldx + c0;
            const float* mu = E.spade_mean + b * 32 + c0;
            const float* rs = E.spade_rstd + b * 32 + c0;
            const float nz = (E.spade_noise && E.spade_noise_w) ? __ldg(E.spade_noise_w) * __ldg(E.spade_noise + row) : 0.f;
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j4 = 0; j4 < EPI_COLS / 4; ++j4) {
              const float4 xv = __ldg(reinterpret_cast<const float4*>(xr) + j4);
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu) + j4);
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(rs) + j4);
              const float4 bg = __ldg(reinterpret_cast<const float4*>(E.bias + c0) + j4);
              const float4 bb = __ldg(reinterpret_cast<const float4*>(E.bias + 32 + c0) + j4);
              const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w};
              const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, gs[4] = {bg.x, bg.y, bg.z, bg.w}, bs[4] = {bb.x, bb.y, bb.z, bb.w};
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float g = __uint_as_float(rg[j4 * 4 + k]) * E.alpha + gs[k];
                const float bt = __uint_as_float(rb[j4 * 4 + k]) * E.alpha + bs[k];
                const float t = ((xs[k] - ms[k]) * rr[k]) * (1.f + g) + bt + nz;      // same association as spade_pixel()
                o[k] = t > 0.f ? t : 0.2f * t;
              }
              if (E.out_f32)
                *reinterpret_cast<float4*>(E.out_f32 + static_cast<size_t>(row) * E.ld_out + c0 + j4 * 4) =
                    make_float4(o[0], o[1], o[2], o[3]);
              if (E.out_hi) {
                split_bf16x2(o[0], o[1], hw[(j4 & 1) * 2], lw[(j4 & 1) * 2]);
                split_bf16x2(o[2], o[3], hw[(j4 & 1) * 2 + 1], lw[(j4 & 1) * 2 + 1]);
                if (j4 & 1) {      // eight channels ready: one 16-byte store per part
                  const size_t idx = static_cast<size_t>(row) * E.ld_split + c0 + (j4 >> 1) * 8;
                  *reinterpret_cast<uint4*>(E.out_hi + idx) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                  *reinterpret_cast<uint4*>(E.out_lo + idx) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
              }
            }
          }
        }
      } else
      for (int c = half; c <= last_c; c += EPI_WARPS / 4) {
        uint32_t r[EPI_COLS];
        tmem_ld_32x16(taddr + c * EPI_COLS, r);
        tmem_ld_wait();
        if (c == last_c && grp == ngrp - 1) {  // accumulator(s) fully read by this warp: release before the global traffic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
        }
        if (P.debug & 1) continue;
        if (ROWSTAT) {
          // this thread holds columns nb … nb+15 of its own row: fold them into the running statistics
          const int nb = n0 + c * EPI_COLS;
          float v[EPI_COLS];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < EPI_COLS; ++j) {
            const int col = nb + j;
            float x = __uint_as_float(r[j]) * P.epi.alpha;
            if (col < P.epi.rowstat_cols) {
              if (P.epi.bias) x += __ldg(P.epi.bias + col);
            } else {
              x = -INFINITY;
            }
            v[j] = x;
            cm = fmaxf(cm, x);
          }
          if (cm > rs_m) {           // strictly greater: an equal value later in the row never replaces the first index
            int a = 0;
#pragma unroll
            for (int j = EPI_COLS - 1; j >= 0; --j) a = (v[j] == cm) ? j : a;
            rs_s *= expf(rs_m - cm);      // exp(−inf) = 0 on the first chunk
            rs_m = cm;
            rs_arg = nb + a;
          }
          if (rs_m > -INFINITY) {
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) rs_s += expf(v[j] - rs_m);
          }
          continue;
        }
        const int wsw = (lane >> 1) & 3;
#pragma unroll
        for (int g = 0; g < EPI_COLS / 4; ++g)
          *reinterpret_cast<float4*>(stage + lane * EPI_COLS + ((g ^ wsw) << 2)) =
              make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                          __uint_as_float(r[g * 4 + 3]));
        if (P.tma_out) {
          // plain fp32 output: the staged 32 × 16 block (64-byte-swizzle layout) leaves as one bulk tensor store
          if (P.epi.bias) {      // bias-only epilogue: added to the thread's own staged row (the accumulator registers
                                 // are dead by now, so the hot variants keep their register budget)
            const int nb = n0 + c * EPI_COLS;
#pragma unroll 1
            for (int g = 0; g < EPI_COLS / 4; ++g) {
              if (nb + g * 4 >= P.N) break;
              float4* sp = reinterpret_cast<float4*>(stage + lane * EPI_COLS + ((g ^ wsw) << 2));
              const float4 bv = __ldg(reinterpret_cast<const float4*>(P.epi.bias + nb) + g);
              float4 v = *sp;
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
              *sp = v;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapOut, smem_u32(stage), n0 + c * EPI_COLS, m0 + q * 32);
            tma_store_commit();
            tma_store_wait_read();
          }
          __syncwarp();
          continue;
        }
        __syncwarp();
        const int n = n0 + c * EPI_COLS + cq * 4;
        const bool col_ok = n < P.N;
        if (EPI >= 6) {       // compile-time epilogue bodies (see lean_chunk)
          if (col_ok) lean_chunk<EPI>(P, stage, m0 + q * 32, n, sub, cq);
          __syncwarp();
          continue;
        }
        if (P.splits > 1) {   // raw partial sums; the reduce kernel finishes the job
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + sub, row = m0 + q * 32 + rr;
            const float4 acc = *reinterpret_cast<const float4*>(stage + rr * EPI_COLS + ((cq ^ ((rr >> 1) & 3)) << 2));
            if (row < P.M && col_ok)
              *reinterpret_cast<float4*>(P.part + (static_cast<size_t>(split) * P.M + row) * P.N + n) = acc;
          }
        } else if (ACT) {     // activation epilogues: arithmetic-bound, kept rolled (small code)
#pragma unroll 1
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + sub, row = m0 + q * 32 + rr;
            const float4 acc = *reinterpret_cast<const float4*>(stage + rr * EPI_COLS + ((cq ^ ((rr >> 1) & 3)) << 2));
            if (row < P.M && col_ok) {
              EpiIn in;
              epilogue_fetch(P, row, n, in);
              epilogue_vec4<true>(P, row, n, acc, in);
            }
          }
        } else {              // load-bound epilogues: all global inputs of the chunk in flight before any arithmetic
          EpiIn in[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = m0 + q * 32 + it * 8 + sub;
            if (row < P.M && col_ok) epilogue_fetch(P, row, n, in[it]);
          }
          float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + sub, row = m0 + q * 32 + rr;
            const float4 acc = *reinterpret_cast<const float4*>(stage + rr * EPI_COLS + ((cq ^ ((rr >> 1) & 3)) << 2));
            if (row < P.M && col_ok) {
              const float4 v = epilogue_vec4<false>(P, row, n, acc, in[it]);
              cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
            }
          }
          if (P.epi.colsum_part) {   // sums over this warp's 32 rows: fold the 8 row lanes, lanes 0-3 write 16 columns
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            }
            if (sub == 0 && col_ok && m0 + q * 32 < P.M)
              *reinterpret_cast<float4*>(P.epi.colsum_part + static_cast<size_t>((m0 >> 5) + q) * P.N + n) = cs;
          }
        }
        __syncwarp();
      }
      }   // grp
      if (FOLD && P.splits > 1) {
        // Split-K without a second kernel: when every epilogue warp of this CTA has stored its partial sums of
        // (tile, split), one thread counts the arrival; the CTA that arrives last re-reads all the tile's partials
        // (L2-resident, written moments ago) in split order and writes the final values.
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (e == 0 && lane == 0) {
          const int t = item % P.tiles_per_split;
          int* ctr = P.tile_counter + (P.cluster == 2 ? 2 * t + crank : t);
          const int old = atomicAdd(ctr, 1);
          const int last = (old == P.splits - 1);
          if (last) *ctr = 0;                      // ready for the next launch on this workspace
          splitk_last = last;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (*reinterpret_cast<volatile int*>(&splitk_last)) {
          __threadfence();
          const size_t mn = static_cast<size_t>(P.M) * P.N;
          for (int c = half; c <= last_c; c += EPI_WARPS / 4) {
            const int n = n0 + c * EPI_COLS + cq * 4;
            if (n >= P.N) continue;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int row = m0 + q * 32 + it * 8 + sub;
              if (row >= P.M) continue;
              const float* src = P.part + static_cast<size_t>(row) * P.N + n;
              float4 a = __ldcg(reinterpret_cast<const float4*>(src));
              for (int sp = 1; sp < P.splits; ++sp) {
                const float4 b = __ldcg(reinterpret_cast<const float4*>(src + sp * mn));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
              }
              float4* dst = reinterpret_cast<float4*>(P.epi.out_f32 + static_cast<size_t>(row) * P.epi.ld_out + n);
              if (P.epi.flags & EPI_ACCUM) {
                const float4 o = *dst;
                a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
              }
              *dst = a;
            }
          }
        }
      }
      if (ROWSTAT) {
        const int row = m0 + q * 32 + lane;
        if (row < P.M) {
          const int slots = P.tiles_n * (EPI_WARPS / 4);
          float* o = P.epi.rowstat + (static_cast<size_t>(row) * slots + (n0 / P.BN) * (EPI_WARPS / 4) + half) * 3;
          o[0] = rs_m; o[1] = rs_s; o[2] = __int_as_float(rs_arg);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (P.cluster == 2) cluster_sync_all();     // nobody exits while the peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// out[m, n] (+)= Σ_s part[s][m][n]  — second phase of a split-K GEMM (plain fp32 output only)
__global__ void splitk_reduce_kernel(const float4* __restrict__ part, int splits, size_t mn4, int n4, float* out,
                                     int ld_out, int accumulate) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < mn4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 a = __ldg(part + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldg(part + s * mn4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    // the partial-sum workspace holds < 2^32 float4s: 32-bit division (the 64-bit one costs ~10× more)
    const unsigned i32 = static_cast<unsigned>(i), m = i32 / static_cast<unsigned>(n4), c = i32 - m * n4;
    float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(m) * ld_out) + c;
    if (accumulate) {
      const float4 o = *dst;
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
    *dst = a;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box0, box1, swz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box0 == o.box0 &&
           box1 == o.box1 && swz == o.swz;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box0); mix(k.box1); mix(k.swz);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;
std::atomic<long long> g_launches{0};

// optional per-launch timing (bench.py roofline): events bracket every GEMM on its own stream
struct TimedLaunch { cudaEvent_t e0, e1; double flops; int M, N, K, passes, a_mn, b_mn, flags, grid; };
bool g_timing = false;
std::vector<TimedLaunch> g_timed;

// bf16 row-major [outer, inner] with leading dimension ld (elements); box = box0 (inner) × box1 (outer).
int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box0,
             uint32_t box1, CUtensorMapSwizzle swz) {
  MapKey key{ptr, inner, outer, ld, box0, box1, static_cast<uint32_t>(swz)};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -11;
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(key, *out);
  return 0;
}

// Weight matrix [rows, kext] (K index = tap·C + c) seen as rank 3 (c, row, dx) with the dx axis striding by C elements:
// one box (32, box_rows, 3) lands in shared memory as three consecutive [box_rows, 32] tap matrices.
int make_map_taps(CUtensorMap* out, const void* ptr, uint64_t kext, uint64_t rows, uint64_t ld, uint64_t tap_stride,
                  uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t gdim[3] = {kext, rows, 3};
  cuuint64_t gstride[2] = {ld * 2, tap_stride * 2};
  cuuint32_t box[3] = {32, box_rows, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

// NHWC bf16 tensor [B, H, W, C] as a rank-4 tensor map (C innermost) with box (bc, bw, bh, bb).
struct Map4Key {
  const void* ptr;
  int B, H, W, C, bc, bw, bh, bb;
  bool operator==(const Map4Key& o) const {
    return ptr == o.ptr && B == o.B && H == o.H && W == o.W && C == o.C && bc == o.bc && bw == o.bw && bh == o.bh &&
           bb == o.bb;
  }
};
struct Map4KeyHash {
  size_t operator()(const Map4Key& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.B); mix(k.H); mix(k.W); mix(k.C); mix(k.bc); mix(k.bw); mix(k.bh); mix(k.bb);
    return h;
  }
};
std::unordered_map<Map4Key, CUtensorMap, Map4KeyHash> g_maps4;

int make_map4(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int bc, int bw, int bh, int bb,
              CUtensorMapSwizzle swz) {
  Map4Key key{ptr, B, H, W, C, bc, bw, bh, bb};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps4.find(key);
    if (it != g_maps4.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(bc), static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                       static_cast<cuuint32_t>(bb)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -11;
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps4.size() > 4096) g_maps4.clear();
  g_maps4.emplace(key, *out);
  return 0;
}

// fp32 row-major [outer, inner] output matrix, box = box0 × box1, 64-byte swizzle (the epilogue staging layout)
int make_map_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box0,
                 uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int BK, int A_MN, int B_MN, int NPARTS, int EPI>
int launch_variant(int grid, size_t smem, cudaStream_t stream, const CUtensorMap& mAhi, const CUtensorMap& mAlo,
                   const CUtensorMap& mBhi, const CUtensorMap& mBlo, const CUtensorMap& mOut, const KParams& P) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BK, A_MN, B_MN, NPARTS, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  static const int force_cl = env_int("XLX_GEMM_FORCE_CLUSTER_LAUNCH", 0);    // experiment: cluster launch, independent CTAs
  attr[0].val.clusterDim.x = (force_cl && grid % 2 == 0) ? 2 : P.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<BK, A_MN, B_MN, NPARTS, EPI>, mAhi, mAlo, mBhi, mBlo, mOut, P);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

template <int BK>
int launch_bk(const GemmProblem& p, cudaStream_t stream) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  KParams P;
  memset(&P, 0, sizeof(P));
  P.M = p.M; P.N = p.N; P.K = p.K;
  int BN = 256;
  if (p.N <= 64) BN = 64; else if (p.N <= 128) BN = 128;
  // narrow convolutions (32 → 32, ToRGB): an N = 64 instruction would spend half of its tensor-pipe time on padding
  static const int conv_bn32 = env_int("XLX_CONV_BN32", 0);
  if (conv_bn32 && p.conv.enabled && p.N <= 32) BN = 32;
  // Small-M GEMMs (inference at sampling batch sizes: M = 2048 rows → 48 tiles of 128 × 256 on 148 SMs): narrower tiles
  // put more SMs to work; one output element's accumulation order does not depend on the tile width, so results are
  // bit-identical.  Only when the wide tiling leaves more than 40 % of the SMs without a tile.
  static const int narrow_on = env_int("XLX_GEMM_NARROW_SMALL_M", 1);
  if (narrow_on && !p.conv.enabled && !p.epi.rowstat && !p.splitk_ws) {
    const int tm = (p.M + BM - 1) / BM;
    while (BN > 64 && tm * ((p.N + BN - 1) / BN) * 10 < num_sms * 6) BN >>= 1;
  }
  int force_bn = env_int("XLX_GEMM_BN", 0);
  if (force_bn) BN = force_bn;
  P.BN = BN;
  P.nparts = (p.passes == 3) ? 2 : 1;
  P.a_mn = p.a.mn_major; P.b_mn = p.b.mn_major;
  P.a_part_bytes = BM * BK * 2;
  P.b_part_bytes = BN * BK * 2;
  P.stage_bytes = P.nparts * (P.a_part_bytes + P.b_part_bytes);
  P.stage_tx = P.stage_bytes;
  int stages = (SMEM_LIMIT - 2048 - 1024 - STAGE_BYTES) / static_cast<int>(P.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return -3;
  P.num_stages = stages;
  P.tiles_m = (p.M + BM - 1) / BM;
  P.tiles_n = (p.N + BN - 1) / BN;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * BN)) cols <<= 1;
  P.tmem_cols = cols;
  P.epi = p.epi;
  static const int debug = env_int("XLX_GEMM_DEBUG", 0);
  P.debug = debug;

  CUtensorMap mAhi, mAlo, mBhi, mBlo;
  const CUtensorMapSwizzle swzK = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  auto mk = [&](CUtensorMap* m, const GemmOperand& o, const __nv_bfloat16* ptr, int rows, int box_rows) -> int {
    const int kext = o.kext ? o.kext : p.K;
    if (!o.mn_major) return make_map(m, ptr, kext, rows, o.ld, BK, box_rows, swzK);
    return make_map(m, ptr, rows, kext, o.ld, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  const int a_rows = p.a.rows ? p.a.rows : p.M, b_rows = p.b.rows ? p.b.rows : p.N;
  P.conv = 0;
  if (p.conv.enabled) {
    const ConvGeometry& g = p.conv;
    const int hw = g.H * g.W;
    if (g.H < 1 || g.W < 1 || (g.taps != 9 && g.taps != 1) || g.C % BK || p.K != g.taps * g.C || p.M % hw ||
        p.a.mn_major)
      return -6;
    int bw, bh, bb;
    if (hw >= BM) {
      if (hw % BM) return -6;
      if (g.W >= BM) { if (g.W % BM) return -6; bw = BM; bh = 1; }
      else { if (BM % g.W) return -6; bw = g.W; bh = BM / g.W; }
      bb = 1;
    } else {
      if (BM % hw) return -6;
      bw = g.W; bh = g.H; bb = BM / hw;
    }
    const int nimg = p.M / hw;
    static const int halo_on = env_int("XLX_CONV_HALO", 1);
    const bool halo = halo_on && BK == 32 && g.taps == 9 && g.W % BM == 0 && !p.b.mn_major;
    if (halo) { bw = BM + 2; bh = 1; bb = 1; }
    if ((rc = make_map4(&mAhi, p.a.hi, nimg, g.H, g.W, g.C, BK, bw, bh, bb, swzK))) return rc;
    if (P.nparts == 2) { if ((rc = make_map4(&mAlo, p.a.lo, nimg, g.H, g.W, g.C, BK, bw, bh, bb, swzK))) return rc; }
    else mAlo = mAhi;
    P.conv = halo ? 2 : 1; P.conv_H = g.H; P.conv_W = g.W; P.conv_taps = g.taps; P.conv_kb_per_tap = g.C / BK;
    P.conv_G = 1;
    if (halo) {
      const uint32_t a_box = (BM + 2) * BK * 2;                       // 130 pixel rows of 64 bytes
      P.a_part_bytes = (a_box + 1023u) & ~1023u;                      // parts stay 1024-byte aligned
      P.b_part_bytes = 3 * BN * BK * 2;                               // the three dx taps of this filter row
      P.stage_bytes = P.nparts * (P.a_part_bytes + P.b_part_bytes);
      P.stage_tx = P.nparts * (a_box + P.b_part_bytes);
      stages = (SMEM_LIMIT - 2048 - 1024 - STAGE_BYTES) / static_cast<int>(P.stage_bytes);
      if (stages > MAX_STAGES) stages = MAX_STAGES;
      if (stages < 2) return -3;
      P.num_stages = stages;
    }
  } else {
    if ((rc = mk(&mAhi, p.a, p.a.hi, a_rows, BM))) return rc;
    if (P.nparts == 2) { if ((rc = mk(&mAlo, p.a, p.a.lo, a_rows, BM))) return rc; }
    else mAlo = mAhi;
  }
  // CTA pairs (cluster of 2 along M) share the B tile through TMA multicast: a third less operand traffic per CTA.
  static const int cluster_on = env_int("XLX_GEMM_CLUSTER", 1);
  P.cluster = (cluster_on && !P.conv && BN >= 128 && P.tiles_m >= 2 && num_sms >= 2) ? 2 : 1;
  const int b_box_rows = P.cluster == 2 ? BN / 2 : BN;
  if ((rc = mk(&mBhi, p.b, p.b.hi, b_rows, b_box_rows))) return rc;
  if (P.nparts == 2) { if ((rc = mk(&mBlo, p.b, p.b.lo, b_rows, b_box_rows))) return rc; }
  else mBlo = mBhi;
  P.conv_tapbox = 0;
  static const int tapbox_on = env_int("XLX_CONV_TAPBOX", 1);
  if (P.conv == 2 && tapbox_on) {
    const uint64_t kext = p.b.kext ? p.b.kext : p.K, C = static_cast<uint64_t>(p.conv.C);
    CUtensorMap t_hi, t_lo;
    int trc = make_map_taps(&t_hi, p.b.hi, kext, b_rows, p.b.ld, C, BN, swzK);
    if (!trc && P.nparts == 2) trc = make_map_taps(&t_lo, p.b.lo, kext, b_rows, p.b.ld, C, BN, swzK);
    if (!trc) {       // (the driver rejects the non-monotonic strides → keep the three 2-D requests)
      mBhi = t_hi;
      mBlo = P.nparts == 2 ? t_lo : t_hi;
      P.conv_tapbox = 1;
    }
  }
  // row-group mode: G tiles per item share the weight groups and the haloed rows (needs the rank-3 weight box, a
  // single n-tile and a plain enough epilogue that the G accumulators can be drained one after the other)
  static const int rowgroup_on = env_int("XLX_CONV_ROWGROUP", 1);
  if (P.conv == 2 && P.conv_tapbox && rowgroup_on && P.tiles_n == 1 && BN <= 64 && p.conv.H % 4 == 0 && !p.epi.rowstat &&
      !p.epi.colsum_part) {
    P.conv = 3;
    P.conv_G = 4;
    P.b_group_bytes = 3 * BN * BK * 2;
    P.b_slots = 4;
    P.stage_bytes = P.nparts * P.a_part_bytes;                    // the A ring carries the haloed rows only
    const int fixed = 2048 + 1024 + STAGE_BYTES + P.b_slots * P.nparts * static_cast<int>(P.b_group_bytes);
    stages = (SMEM_LIMIT - fixed) / static_cast<int>(P.stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 3) { P.conv = 2; P.conv_G = 1; }                  // does not fit: keep the one-tile halo mode
    else {
      P.num_stages = stages;
      P.b_ring_off = stages * P.stage_bytes;
      uint32_t cols = 32;
      while (cols < static_cast<uint32_t>(2 * P.conv_G * BN)) cols <<= 1;
      P.tmem_cols = cols;
    }
    if (P.conv == 2) {   // restore the halo-mode stage geometry
      P.stage_bytes = P.nparts * (P.a_part_bytes + 3 * BN * BK * 2);
      stages = (SMEM_LIMIT - 2048 - 1024 - STAGE_BYTES) / static_cast<int>(P.stage_bytes);
      if (stages > MAX_STAGES) stages = MAX_STAGES;
      P.num_stages = stages;
    }
  }
  // tiles per split: single tiles, or pairs of m-tiles (an odd last m-tile gets an idle partner); row-group
  // convolution: groups of conv_G vertically adjacent tiles
  const int num_tiles = P.conv == 3 ? P.tiles_m / P.conv_G
                        : P.cluster == 2 ? ((P.tiles_m + 1) / 2) * P.tiles_n : P.tiles_m * P.tiles_n;
  const int cta_per_item = P.cluster;
  const int nkb = (P.conv == 2) ? 3 * P.conv_kb_per_tap : (p.K + BK - 1) / BK;
  // split-K: only for plain fp32-output GEMMs (weight gradients) whose tile count leaves most SMs idle
  P.splits = 1; P.kb_per_split = nkb; P.part = nullptr;
  const bool plain = p.epi.out_f32 && !p.epi.bias && !p.epi.addend && !p.epi.addend_hi && !p.epi.out_hi &&
                     !p.epi.out_u && !(p.epi.flags & ~EPI_ACCUM) && p.epi.alpha == 1.0f && !p.epi.drop.threshold &&
                     !p.epi.spade_x;
  static const int splitk_on = env_int("XLX_GEMM_SPLITK", 1);
  if (splitk_on && plain && p.splitk_ws && num_tiles * cta_per_item < num_sms) {
    // pick the split count whose work items fill whole waves of SMs best (ties → fewer splits)
    const size_t need = static_cast<size_t>(p.M) * p.N;
    int maxS = nkb / 8;                                         // ≥ 8 k-blocks per split
    if (maxS > 16) maxS = 16;
    int best = 1;
    double best_eff = 0.0;
    const size_t ws_data = p.splitk_ws_floats > kSplitkCounters ? p.splitk_ws_floats - kSplitkCounters : 0;
    for (int S = 1; S <= maxS && need * S <= ws_data; ++S) {
      const int items = num_tiles * S * cta_per_item;
      const int waves = (items + num_sms - 1) / num_sms;
      const double eff = static_cast<double>(items) / (static_cast<double>(waves) * num_sms);
      if (eff > best_eff + 0.03) { best_eff = eff; best = S; }
    }
    if (best > 1) {
      P.kb_per_split = (nkb + best - 1) / best;
      P.splits = (nkb + P.kb_per_split - 1) / P.kb_per_split;
      P.part = p.splitk_ws;
      static const int fold_on = env_int("XLX_GEMM_SPLITK_FOLD", 0);
      if (fold_on && P.a_mn && P.b_mn && num_tiles * cta_per_item <= kSplitkCounters)
        P.tile_counter = reinterpret_cast<int*>(p.splitk_ws + (p.splitk_ws_floats - kSplitkCounters));
    }
  }
  // plain fp32 output (no split-K partials): TMA store from the staging buffers
  CUtensorMap mOut = mAhi;
  P.tma_out = 0;
  static const int tma_out_on = env_int("XLX_GEMM_TMA_OUT", 1);
  // (a bias is added in registers before the block is staged: bias-only epilogues — the generator's narrow convolutions,
  // the heads' logits — take this path too; measured on the 32 → 32 convolutions at 256²: their st.global epilogue was
  // the bottleneck, profiles/r02_generator_ablation.txt)
  static const int tma_bias_on = env_int("XLX_GEMM_TMA_OUT_BIAS", 1);
  const bool plain_or_bias = p.epi.out_f32 && !p.epi.addend && !p.epi.addend_hi && !p.epi.out_hi && !p.epi.out_u &&
                             !p.epi.flags && p.epi.alpha == 1.0f && !p.epi.drop.threshold && !p.epi.spade_x &&
                             !p.epi.rowstat && !p.epi.colsum_part && (tma_bias_on || !p.epi.bias) &&
                             !(reinterpret_cast<uintptr_t>(p.epi.bias) & 15);
  if (tma_out_on && (plain || plain_or_bias) && P.splits == 1 && !(p.epi.flags & EPI_ACCUM) &&
      !(reinterpret_cast<uintptr_t>(p.epi.out_f32) & 15)) {
    if ((rc = make_map_f32(&mOut, p.epi.out_f32, p.N, p.M, p.epi.ld_out, EPI_COLS, 32))) return rc;
    P.tma_out = 1;
  }
  const int num_items = num_tiles * P.splits;
  P.tiles_per_split = num_tiles;
  P.num_items = num_items;
  int grid = num_items * cta_per_item < num_sms ? num_items * cta_per_item : num_sms;
  if (P.cluster == 2) grid &= ~1;
  const size_t smem = static_cast<size_t>(P.num_stages) * P.stage_bytes + 1024 + STAGE_BYTES +
                      (P.conv == 3 ? static_cast<size_t>(P.b_slots) * P.nparts * P.b_group_bytes : 0);
  TimedLaunch tl{};
  if (g_timing) {
    cudaEventCreate(&tl.e0);
    cudaEventCreate(&tl.e1);
    tl.flops = 2.0 * p.M * p.N * p.K;
    tl.M = p.M; tl.N = p.N; tl.K = p.K; tl.passes = p.passes; tl.a_mn = p.a.mn_major; tl.b_mn = p.b.mn_major;
    tl.flags = p.epi.flags | (p.epi.out_f32 ? 256 : 0) | (p.epi.out_hi ? 512 : 0) | (p.epi.addend_hi ? 1024 : 0) |
               (p.epi.out_u ? 2048 : 0) | (p.epi.addend ? 4096 : 0);
    tl.grid = grid;
    cudaEventRecord(tl.e0, stream);
  }
  {
    const bool act = (p.epi.flags & (EPI_GELU | EPI_TANH | EPI_RELU | EPI_GELU_GRAD)) != 0 || p.epi.drop.threshold != 0;
    const int v = (P.a_mn ? 4 : 0) | (P.b_mn ? 2 : 0) | (P.nparts == 2 ? 1 : 0);
    int lrc = 0;
#define XLX_LAUNCH(A, B, N)                                                                                   \
  lrc = act ? launch_variant<BK, A, B, N, 1>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P)             \
            : launch_variant<BK, A, B, N, 0>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P)
    // compile-time epilogue bodies for the configurations that carry the training step (see lean_chunk)
    int lean = 0;
    static const int lean_on = env_int("XLX_GEMM_LEAN_EPI", 1);
    if (lean_on && v == 1 && !P.conv && !p.epi.rowstat && !p.epi.spade_x && !P.tile_counter && P.splits == 1 &&
        !P.tma_out && p.epi.alpha == 1.0f && !p.epi.colsum_part && !p.epi.addend &&
        !p.epi.out_u16 && !p.epi.u_in16) {
      const GemmEpilogue& e = p.epi;
      const bool split_out = e.out_hi && e.out_lo && !e.out_f32;
      if (e.drop.threshold) {
        if (e.flags == 0 && e.bias && e.addend_hi && e.addend_lo && e.out_f32 && !e.out_hi && !e.out_u) lean = 10;
      } else
      if (((e.flags == (EPI_GELU | EPI_SAVE_DGELU) && e.out_u) || (e.flags == EPI_GELU && !e.out_u)) && e.bias && split_out &&
          !e.addend_hi)
        lean = 6;          // training (gelu' saved) or inference (nothing saved) FFN-1 forward
      else if (e.flags == 0 && split_out && !e.out_u && !e.addend_hi) lean = 7;
      else if (e.flags == 0 && e.addend_hi && e.addend_lo && e.out_f32 && !e.out_hi && !e.out_u) lean = 8;   // bias optional
      else if (e.flags == EPI_MUL && e.u_in && split_out && !e.out_u && !e.addend_hi && !e.bias) lean = 9;
    }
    if (lean) {
      if constexpr (BK == 32) {
        if (lean == 6) lrc = launch_variant<BK, 0, 0, 2, 6>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
        else if (lean == 7) lrc = launch_variant<BK, 0, 0, 2, 7>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
        else if (lean == 8) lrc = launch_variant<BK, 0, 0, 2, 8>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
        else if (lean == 10) lrc = launch_variant<BK, 0, 0, 2, 10>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
        else lrc = launch_variant<BK, 0, 0, 2, 9>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
      } else {
        lrc = -1;
      }
    } else if (p.epi.spade_x) {                // γ/β convolution of the generator (checked by gemm_launch)
      lrc = P.nparts == 2 ? launch_variant<BK, 0, 0, 2, 4>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P)
                          : launch_variant<BK, 0, 0, 1, 4>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
    } else if (p.epi.rowstat) {                // K-major operands only (checked by gemm_launch)
      lrc = P.nparts == 2 ? launch_variant<BK, 0, 0, 2, 2>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P)
                          : launch_variant<BK, 0, 0, 1, 2>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
    } else if (P.tile_counter) {               // folded split-K: weight-gradient layout only (see below)
      lrc = P.nparts == 2 ? launch_variant<BK, 1, 1, 2, 3>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P)
                          : launch_variant<BK, 1, 1, 1, 3>(grid, smem, stream, mAhi, mAlo, mBhi, mBlo, mOut, P);
    } else
    switch (v) {
      case 0: XLX_LAUNCH(0, 0, 1); break;
      case 1: XLX_LAUNCH(0, 0, 2); break;
      case 2: XLX_LAUNCH(0, 1, 1); break;
      case 3: XLX_LAUNCH(0, 1, 2); break;
      case 4: XLX_LAUNCH(1, 0, 1); break;
      case 5: XLX_LAUNCH(1, 0, 2); break;
      case 6: XLX_LAUNCH(1, 1, 1); break;
      default: XLX_LAUNCH(1, 1, 2); break;
    }
#undef XLX_LAUNCH
    if (lrc) return lrc;
  }
  if (P.splits > 1 && !P.tile_counter) {
    const size_t mn4 = static_cast<size_t>(p.M) * p.N / 4;
    size_t blocks = (mn4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream,
               reinterpret_cast<const float4*>(P.part), P.splits, mn4, p.N / 4, p.epi.out_f32, p.epi.ld_out,
               (p.epi.flags & EPI_ACCUM) ? 1 : 0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (g_timing) {
    cudaEventRecord(tl.e1, stream);
    g_timed.push_back(tl);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace

long long gemm_launch_count() { return g_launches.load(); }

int tma_map_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer, int swizzle_bytes) {
  const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return make_map(out, ptr, inner, outer, ld, box_inner, box_outer, swz);
}
int gemm_rowstat_slots(int N) {
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);      // launch_bk's tile width (XLX_GEMM_BN must not be forced)
  return ((N + BN - 1) / BN) * (EPI_WARPS / 4);
}
// splits · tiles ≤ 2 · #SMs and every tile is ≤ 128 × 256 outputs
size_t gemm_splitk_ws_floats() { return static_cast<size_t>(2 * 148) * BM * 256 + kSplitkCounters; }
int gemm_splitk_ws_reset(float* splitk_ws, cudaStream_t stream) {
  if (!splitk_ws) return 0;
  cudaError_t e = cudaMemsetAsync(splitk_ws + (gemm_splitk_ws_floats() - kSplitkCounters), 0, kSplitkCounters * 4, stream);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

bool gemm_timing_active() { return g_timing; }
void gemm_timing_begin() {
  g_timed.clear();
  g_timing = true;
  g_pdl_suspended = true;
}
int gemm_timing_end(double* total_ms, double* total_flops, long long* launches) {
  g_timing = false;
  g_pdl_suspended = false;
  double ms = 0, fl = 0;
  FILE* log = nullptr;
  if (const char* path = getenv("XLX_GEMM_LOG")) log = fopen(path, "w");
  if (log) fprintf(log, "M,N,K,passes,a_mn,b_mn,epi,grid,us\n");
  for (auto& t : g_timed) {
    cudaError_t e = cudaEventSynchronize(t.e1);
    if (e != cudaSuccess) return static_cast<int>(e);
    float m = 0;
    cudaEventElapsedTime(&m, t.e0, t.e1);
    ms += m; fl += t.flops;
    if (log) fprintf(log, "%d,%d,%d,%d,%d,%d,%d,%d,%.2f\n", t.M, t.N, t.K, t.passes, t.a_mn, t.b_mn, t.flags, t.grid, m * 1e3);
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  if (log) fclose(log);
  *total_ms = ms; *total_flops = fl; *launches = static_cast<long long>(g_timed.size());
  g_timed.clear();
  return 0;
}

int gemm_launch(const GemmProblem& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return -1;
  if (p.passes != 1 && p.passes != 3) return -1;
  if (!p.a.hi || !p.b.hi) return -1;
  if (p.passes == 3 && (!p.a.lo || !p.b.lo)) return -1;
  if ((!p.conv.enabled && (p.a.ld % 8)) || (p.b.ld % 8) || (p.N % 4)) return -2;  // TMA: 16-byte global strides; epilogue: 4-column vectors
  if ((reinterpret_cast<uintptr_t>(p.a.hi) | reinterpret_cast<uintptr_t>(p.b.hi) |
       reinterpret_cast<uintptr_t>(p.a.lo) | reinterpret_cast<uintptr_t>(p.b.lo)) & 15)
    return -2;
  const GemmEpilogue& E = p.epi;
  if ((E.out_f32 && (E.ld_out % 4)) || (E.out_hi && (E.ld_split % 4)) ||
      ((E.addend || E.addend_hi) && (E.ld_addend % 4)) ||
      ((E.out_u || E.u_in || E.out_u16 || E.u_in16) && (E.ld_u % 4)))
    return -2;
  if ((E.flags & EPI_GELU_GRAD) && !E.u_in) return -1;
  if (E.colsum_part && ((E.flags & (EPI_GELU | EPI_TANH | EPI_RELU | EPI_GELU_GRAD)) || p.splitk_ws || E.drop.threshold))
    return -1;
  if (E.drop.threshold && (E.rowstat || p.conv.enabled)) return -1;
  if (E.spade_x) {
    if (p.a.mn_major || p.b.mn_major || p.N != 64 || !E.bias || !E.spade_mean || !E.spade_rstd || E.flags || E.rowstat ||
        E.addend || E.addend_hi || E.out_u || E.colsum_part || E.drop.threshold || p.splitk_ws || (!E.out_f32 && !E.out_hi) ||
        (E.out_hi && !E.out_lo) || (E.spade_ldx % 4) || (E.out_f32 && (E.ld_out % 4)) || (E.out_hi && (E.ld_split % 8)) ||
        getenv("XLX_GEMM_BN"))
      return -1;
  }
  if ((E.flags & EPI_MUL) && !E.u_in && !E.u_in16) return -1;
  if (E.addend_hi && !E.addend_lo) return -1;
  if (E.rowstat && (p.a.mn_major || p.b.mn_major || E.out_f32 || E.out_hi || E.out_u || E.out_u16 || E.colsum_part || p.splitk_ws || E.flags ||
                    E.rowstat_cols < 1 || E.rowstat_cols > p.N || getenv("XLX_GEMM_BN")))
    return -1;
  return launch_bk<32>(p, stream);
}

}  // namespace xlx
