// Attention core of LxmertAttention.forward (HF modeling_lxmert.py:238-274) and its backward on tensor cores.
//
// One CTA (4 warps) per (sample, head); sequences are tiny (≤ 64 queries × ≤ 64 keys × 64 features), so the whole
// problem lives in shared memory as split-bf16 tiles and every product runs as warp-level bf16 MMAs
// (mma.sync.m16n8k16) in the same 3-pass split scheme as the big GEMMs (A_lo·B_hi + A_hi·B_lo + A_hi·B_hi, fp32
// accumulate) — fp32-class accuracy at tensor-core speed; the softmax stays in fp32 registers between the two products
// (the score tile never touches shared or global memory in the forward).  Transposed operands (Pᵀ, dSᵀ, and the
// [key][feature] tiles used as K-major B operands) are read with ldmatrix.trans instead of being transposed in memory.
// The kernels are bound by their global loads/stores, not by math, so the operand tiles (Q, K, V, dO — already split
// bf16, written that way by the producing GEMM epilogues) arrive by TMA (one 64×64 box per tile, 128-byte hardware
// swizzle, all boxes of a CTA in flight at once) and several CTAs are resident per SM.
#include <cuda.h>

#include "gemm_sm100.cuh"
#include "kernels.cuh"
#include "xlx_ptx.cuh"

namespace xlx {

namespace {

constexpr int AT = 128;      // threads per CTA
constexpr int TT = 64 * 64;  // elements of a TMA-written tile: 64 rows × 128 B, 128-byte swizzle, 1024-byte aligned

// element (row r, column k) of a TMA tile: the 16-byte chunk index is XOR-ed with the row index modulo 8
__device__ __forceinline__ int swz(int r, int k) { return r * 64 + ((((k >> 3) ^ (r & 7)) << 3) | (k & 7)); }

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ah + al)·(bh + bl) without the lo·lo term, smallest terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(c, al, bh0, bh1);
  mma16816(c, ah, bl0, bl1);
  mma16816(c, ah, bh0, bh1);
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows r0.., 16 columns k0..) of a swizzled row-major tile X[row][k]
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* X, int r0, int k0, int g, int t) {
  a[0] = lds32(X + swz(r0 + g, k0 + 2 * t));
  a[1] = lds32(X + swz(r0 + g + 8, k0 + 2 * t));
  a[2] = lds32(X + swz(r0 + g, k0 + 2 * t + 8));
  a[3] = lds32(X + swz(r0 + g + 8, k0 + 2 * t + 8));
}
// B fragment (n-tile n0, 16 columns k0..) of a swizzled tile Y[n][k]
__device__ __forceinline__ void load_b(uint32_t& b0, uint32_t& b1, const bf16* Y, int n0, int k0, int g, int t) {
  b0 = lds32(Y + swz(n0 + g, k0 + 2 * t));
  b1 = lds32(Y + swz(n0 + g, k0 + 2 * t + 8));
}
// A fragment of the TRANSPOSE of a swizzled tile S[k][m]: rows m0.., columns k0..
__device__ __forceinline__ void load_a_trans(uint32_t (&a)[4], const bf16* S, int m0, int k0, int lane) {
  const int mat = lane >> 3, i = lane & 7;
  const bf16* p = S + swz(k0 + (mat >> 1) * 8 + i, m0 + (mat & 1) * 8);
  ldsm_x4_trans(a, smem_u32(p));
}
// B fragments of two adjacent n-tiles (n0, n0 + 8) from a swizzled tile Z[k][n] (k0..k0+15): r[0..1] → n0, r[2..3] → n0+8
__device__ __forceinline__ void load_b_kmajor(uint32_t (&r)[4], const bf16* Z, int k0, int n0, int lane) {
  const int mat = lane >> 3, i = lane & 7;
  const bf16* p = Z + swz(k0 + (mat & 1) * 8 + i, n0 + (mat >> 1) * 8);
  ldsm_x4_trans(r, smem_u32(p));
}
// Thread 0 arms the barrier and issues one 64×64 box per tile; everybody then waits for the bytes to land.
struct TileSrc { const CUtensorMap* map; int col, row; };
template <int N>
__device__ __forceinline__ void tma_stage(bf16* smem0, uint64_t* bar, const TileSrc (&src)[N]) {
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(smem_u32(bar), N * TT * 2);
#pragma unroll
    for (int i = 0; i < N; ++i) tma_load_2d(smem_u32(smem0 + i * TT), src[i].map, smem_u32(bar), src[i].col, src[i].row);
  }
  __syncthreads();            // barrier initialised before anybody polls it
  mbar_wait(smem_u32(bar), 0);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) { split_bf16x2(x, y, hi, lo); }
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// ---- forward ------------------------------------------------------------------------------------------
// maps: [0] Q hi, [1] Q lo, [2] K hi, [3] K lo, [4] V hi, [5] V lo — 2-D maps over the split [rows, ld] matrices
struct FwdMaps { CUtensorMap m[6]; };
// PACK (self-attention with S ≤ 32 only): one CTA serves G = 64 / S consecutive samples.  Their rows are contiguous
// in memory, so the same 64-row TMA boxes bring them in; the score tile is block-diagonal — entries pairing different
// samples are masked to −inf before the soft-max and contribute exact zeros to P·V.  The language stack's (S ≤ 20)
// kernels were bound by per-CTA latency, not work: a third of the CTAs do the same job.
template <bool DROP, bool PACK>   // compile-time: the inference / p = 0 instantiation carries no dropout code or registers
__global__ void __launch_bounds__(AT)
attn_fwd_mma_kernel(const __grid_constant__ FwdMaps maps, int qcol, int kcol, int vcol,
                    const float* __restrict__ mask, int heads, int Sq, int Sk, bf16* ctx_hi, bf16* ctx_lo,
                    float* ctx_f32, int ld_ctx, float* probs, const DropSite drop, int nbatch) {
  extern __shared__ uint8_t smem_attn_raw[];
  __shared__ __align__(8) uint64_t bar;
  bf16* Qh = reinterpret_cast<bf16*>(smem_attn_raw + ((1024u - (smem_u32(smem_attn_raw) & 1023u)) & 1023u));
  bf16* Ql = Qh + TT;
  bf16* Kh = Ql + TT;
  bf16* Kl = Kh + TT;
  bf16* Vh = Kl + TT;
  bf16* Vl = Vh + TT;
  // S1: per-sample sequence length; packed mode turns (Sq, Sk) into the extent of the G samples this CTA serves
  const int S1 = Sq;
  const int h = blockIdx.x, b = PACK ? blockIdx.y * (64 / S1) : blockIdx.y;
  if (PACK) { const int gc = min(64 / S1, nbatch - b); Sq = Sk = gc * S1; }
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t qrow0 = static_cast<size_t>(b) * S1;
  {
    const int qr = b * S1, kr = PACK ? b * S1 : b * Sk;
    const TileSrc src[6] = {{&maps.m[0], qcol + h * 64, qr}, {&maps.m[1], qcol + h * 64, qr},
                            {&maps.m[2], kcol + h * 64, kr}, {&maps.m[3], kcol + h * 64, kr},
                            {&maps.m[4], vcol + h * 64, kr}, {&maps.m[5], vcol + h * 64, kr}};
    tma_stage<6>(Qh, &bar, src);
  }
  const int r0 = warp * 16;
  if (r0 >= Sq) return;
  const int nk16 = (Sk + 15) >> 4;     // 16-key steps; score n-tiles = 2·nk16
  const int i0 = r0 + g, i1 = r0 + g + 8;
  // packed mode: sample slot and in-sample index of this thread's two query rows
  const int gi0 = PACK ? i0 / S1 : 0, gi1 = PACK ? i1 / S1 : 0;

  // S = Q·Kᵀ
  float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t ah[4], al[4];
    load_a(ah, Qh, r0, kk * 16, g, t);
    load_a(al, Ql, r0, kk * 16, g, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * nk16) {
        uint32_t bh0, bh1, bl0, bl1;
        load_b(bh0, bh1, Kh, nt * 8, kk * 16, g, t);
        load_b(bl0, bl1, Kl, nt * 8, kk * 16, g, t);
        mma3(s[nt], ah, al, bh0, bh1, bl0, bl1);
      }
    }
  }
  // softmax(S/8 + mask) over keys: thread holds rows g, g+8 and columns nt·8 + 2t, +1
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < 2 * nk16) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        if (PACK) {
          const int gj = j / S1;
          const float mk = (j < Sk) ? (mask ? __ldg(mask + static_cast<size_t>(b + gj) * S1 + (j - gj * S1)) : 0.f) : -INFINITY;
          s[nt][e] = (gj == gi0) ? s[nt][e] * 0.125f + mk : -INFINITY;         // other samples' keys: not part of this row
          s[nt][2 + e] = (gj == gi1) ? s[nt][2 + e] * 0.125f + mk : -INFINITY;
        } else {
          const float mk = (j < Sk) ? (mask ? __ldg(mask + static_cast<size_t>(b) * Sk + j) : 0.f) : -INFINITY;
          s[nt][e] = s[nt][e] * 0.125f + mk;          // scores / sqrt(64) then + mask (HF:255-259)
          s[nt][2 + e] = s[nt][2 + e] * 0.125f + mk;
        }
        mx0 = fmaxf(mx0, s[nt][e]);
        mx1 = fmaxf(mx1, s[nt][2 + e]);
      }
    }
  }
  mx0 = quad_max(mx0); mx1 = quad_max(mx1);
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < 2 * nk16) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = expf(s[nt][e] - mx0); sum0 += s[nt][e];
        s[nt][2 + e] = expf(s[nt][2 + e] - mx1); sum1 += s[nt][2 + e];
      }
    }
  }
  const float inv0 = 1.0f / quad_sum(sum0), inv1 = 1.0f / quad_sum(sum1);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < 2 * nk16) {
      s[nt][0] *= inv0; s[nt][1] *= inv0; s[nt][2] *= inv1; s[nt][3] *= inv1;
      if (PACK) {
        // per element: (sample slot, in-sample key) of column j; only same-sample entries exist in probs / get a mask
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e, gj = j / S1, jj = j - gj * S1;
          if (probs && j < Sk) {
            if (i0 < Sq && gj == gi0)
              probs[((static_cast<size_t>(b + gi0) * heads + h) * S1 + (i0 - gi0 * S1)) * S1 + jj] = s[nt][e];
            if (i1 < Sq && gj == gi1)
              probs[((static_cast<size_t>(b + gi1) * heads + h) * S1 + (i1 - gi1 * S1)) * S1 + jj] = s[nt][2 + e];
          }
          if (DROP) {
            if (gj == gi0) s[nt][e] *= drop_prob1(drop, (static_cast<size_t>(b + gi0) * heads + h) * S1 + (i0 - gi0 * S1), S1, jj);
            if (gj == gi1) s[nt][2 + e] *= drop_prob1(drop, (static_cast<size_t>(b + gi1) * heads + h) * S1 + (i1 - gi1 * S1), S1, jj);
          }
        }
      } else {
        if (probs) {
          const int j = nt * 8 + 2 * t;
          float* pr = probs + (static_cast<size_t>(b) * heads + h) * Sq * Sk;
          if (i0 < Sq) { if (j < Sk) pr[i0 * Sk + j] = s[nt][0]; if (j + 1 < Sk) pr[i0 * Sk + j + 1] = s[nt][1]; }
          if (i1 < Sq) { if (j < Sk) pr[i1 * Sk + j] = s[nt][2]; if (j + 1 < Sk) pr[i1 * Sk + j + 1] = s[nt][3]; }
        }
        if (DROP) {   // dropout(attention_probs) (HF:262-266): P·V below sees P ∘ mask / (1 − p)
          const int j = nt * 8 + 2 * t;
          const size_t rbase = (static_cast<size_t>(b) * heads + h) * Sq;
          const float2 m0 = drop_prob2(drop, rbase + i0, Sk, j), m1 = drop_prob2(drop, rbase + i1, Sk, j);
          s[nt][0] *= m0.x; s[nt][1] *= m0.y; s[nt][2] *= m1.x; s[nt][3] *= m1.y;
        }
      }
    }
  }
  // O = P·V: the score accumulators of two adjacent n-tiles are exactly one A fragment of the next product
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk < nk16) {
      uint32_t ah[4], al[4];
      split2(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
      split2(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
      split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
      split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bh[4], bl[4];
        load_b_kmajor(bh, Vh, kk * 16, np * 16, lane);
        load_b_kmajor(bl, Vl, kk * 16, np * 16, lane);
        mma3(o[2 * np], ah, al, bh[0], bh[1], bl[0], bl[1]);
        mma3(o[2 * np + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
      }
    }
  }
  // context rows: through this warp's own 16 rows of the (now dead) Q tiles, then 16-byte row-contiguous stores
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t hi, lo;
    split2(o[nt][0], o[nt][1], hi, lo);
    *reinterpret_cast<uint32_t*>(Qh + (r0 + g) * 64 + nt * 8 + 2 * t) = hi;
    *reinterpret_cast<uint32_t*>(Ql + (r0 + g) * 64 + nt * 8 + 2 * t) = lo;
    split2(o[nt][2], o[nt][3], hi, lo);
    *reinterpret_cast<uint32_t*>(Qh + (r0 + g + 8) * 64 + nt * 8 + 2 * t) = hi;
    *reinterpret_cast<uint32_t*>(Ql + (r0 + g + 8) * 64 + nt * 8 + 2 * t) = lo;
    if (ctx_f32) {
      const int c = h * 64 + nt * 8 + 2 * t;
      if (i0 < Sq) *reinterpret_cast<float2*>(ctx_f32 + (qrow0 + i0) * ld_ctx + c) = make_float2(o[nt][0], o[nt][1]);
      if (i1 < Sq) *reinterpret_cast<float2*>(ctx_f32 + (qrow0 + i1) * ld_ctx + c) = make_float2(o[nt][2], o[nt][3]);
    }
  }
  __syncwarp();
  if (ctx_hi) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 4 + (lane >> 3), ch = lane & 7;     // 4 rows × 8 sixteen-byte chunks per pass
      const int i = r0 + rr;
      if (i < Sq) {
        const size_t idx = (qrow0 + i) * ld_ctx + h * 64 + ch * 8;
        *reinterpret_cast<uint4*>(ctx_hi + idx) = *reinterpret_cast<const uint4*>(Qh + i * 64 + ch * 8);
        if (ctx_lo) *reinterpret_cast<uint4*>(ctx_lo + idx) = *reinterpret_cast<const uint4*>(Ql + i * 64 + ch * 8);
      }
    }
  }
}

// ---- backward -----------------------------------------------------------------------------------------
// dP = dO·Vᵀ;  dS = P ∘ (dP − rowsum(P ∘ dP)) / 8;  dQ = dS·K;  dV = Pᵀ·dO;  dK = dSᵀ·Q.
// maps: Q, K, V, dO (hi, lo each)
struct BwdMaps { CUtensorMap m[8]; };
template <bool DROP, bool PACK>
__global__ void __launch_bounds__(AT)
attn_bwd_mma_kernel(const __grid_constant__ BwdMaps maps, int qcol, int kcol, int vcol, int ocol,
                    const float* __restrict__ probs, int heads, int Sq, int Sk, bf16* dq_hi, bf16* dq_lo, bf16* dk_hi,
                    bf16* dk_lo, bf16* dv_hi, bf16* dv_lo, int ld_d, const DropSite drop, int nbatch) {
  extern __shared__ uint8_t smem_attn_raw[];
  __shared__ __align__(8) uint64_t bar;
  bf16* Qh = reinterpret_cast<bf16*>(smem_attn_raw + ((1024u - (smem_u32(smem_attn_raw) & 1023u)) & 1023u));
  bf16* Ql = Qh + TT;
  bf16* Kh = Ql + TT;
  bf16* Kl = Kh + TT;
  bf16* Vh = Kl + TT;
  bf16* Vl = Vh + TT;
  bf16* Oh = Vl + TT;        // dO
  bf16* Ol = Oh + TT;
  // P and dS ([query][key], same swizzled layout) overlay the V and K tiles once phase 1 no longer needs them
  bf16* Ph = Vh;
  bf16* Pl = Vl;
  bf16* Sh = Kh;
  bf16* Sl = Kl;
  const int S1 = Sq;       // per-sample length; packed mode (see the forward kernel) widens (Sq, Sk) to the CTA's samples
  const int h = blockIdx.x, b = PACK ? blockIdx.y * (64 / S1) : blockIdx.y;
  if (PACK) { const int gc = min(64 / S1, nbatch - b); Sq = Sk = gc * S1; }
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t qrow0 = static_cast<size_t>(b) * S1, krow0 = PACK ? qrow0 : static_cast<size_t>(b) * Sk;
  {
    const int qr = b * S1, kr = PACK ? b * S1 : b * Sk;
    const TileSrc src[8] = {{&maps.m[0], qcol + h * 64, qr}, {&maps.m[1], qcol + h * 64, qr},
                            {&maps.m[2], kcol + h * 64, kr}, {&maps.m[3], kcol + h * 64, kr},
                            {&maps.m[4], vcol + h * 64, kr}, {&maps.m[5], vcol + h * 64, kr},
                            {&maps.m[6], ocol + h * 64, qr}, {&maps.m[7], ocol + h * 64, qr}};
    tma_stage<8>(Qh, &bar, src);
  }
  const int r0 = warp * 16;
  const int nk16 = (Sk + 15) >> 4, nq16 = (Sq + 15) >> 4;
  const int i0 = r0 + g, i1 = r0 + g + 8;

  // phase 1 (warp = 16 query rows): dP = dO·Vᵀ, dS, dQ = dS·K, all in registers
  float dp[8][4], p[8][4];
  float pm[DROP ? 8 : 1][4];      // P ∘ dropout mask
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
    if (DROP) pm[DROP ? nt : 0][0] = pm[DROP ? nt : 0][1] = pm[DROP ? nt : 0][2] = pm[DROP ? nt : 0][3] = 0.f;
  }
  const bool rows_live = r0 < 16 * nq16;
  if (rows_live) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ah[4], al[4];
      load_a(ah, Oh, r0, kk * 16, g, t);
      load_a(al, Ol, r0, kk * 16, g, t);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < 2 * nk16) {
          uint32_t bh0, bh1, bl0, bl1;
          load_b(bh0, bh1, Vh, nt * 8, kk * 16, g, t);
          load_b(bl0, bl1, Vl, nt * 8, kk * 16, g, t);
          mma3(dp[nt], ah, al, bh0, bh1, bl0, bl1);
        }
      }
    }
    float dot0 = 0.f, dot1 = 0.f;
    const float* pr = probs + (static_cast<size_t>(b) * heads + h) * Sq * Sk;
    const int gi0 = PACK ? i0 / S1 : 0, gi1 = PACK ? i1 / S1 : 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * nk16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          if (PACK) {
            // probabilities exist for same-sample pairs only; everything else is an exact zero of the block-diagonal tile
            const int gj = j / S1, jj = j - gj * S1;
            const bool ok0 = i0 < Sq && j < Sk && gj == gi0, ok1 = i1 < Sq && j < Sk && gj == gi1;
            const size_t r0p = (static_cast<size_t>(b + gi0) * heads + h) * S1 + (i0 - gi0 * S1);
            const size_t r1p = (static_cast<size_t>(b + gi1) * heads + h) * S1 + (i1 - gi1 * S1);
            p[nt][e] = ok0 ? __ldg(probs + r0p * S1 + jj) : 0.f;
            p[nt][2 + e] = ok1 ? __ldg(probs + r1p * S1 + jj) : 0.f;
            if (DROP) {
              constexpr int kD = DROP ? 1 : 0;
              const float m0 = ok0 ? drop_prob1(drop, r0p, S1, jj) : 0.f, m1 = ok1 ? drop_prob1(drop, r1p, S1, jj) : 0.f;
              dp[nt][e] *= m0; dp[nt][2 + e] *= m1;
              pm[nt * kD][e] = p[nt][e] * m0; pm[nt * kD][2 + e] = p[nt][2 + e] * m1;
            }
          } else {
            p[nt][e] = (i0 < Sq && j < Sk) ? __ldg(pr + i0 * Sk + j) : 0.f;
            p[nt][2 + e] = (i1 < Sq && j < Sk) ? __ldg(pr + i1 * Sk + j) : 0.f;
          }
        }
        if (DROP && !PACK) {
          // forward: ctx = (P ∘ m)·V with m = mask / (1 − p).  dp holds d(P ∘ m) = dO·Vᵀ → dP = dp ∘ m; the P tile
          // phase 2 contracts with dO (dV = (P ∘ m)ᵀ·dO) is the dropped one: pm below.
          const int j = nt * 8 + 2 * t;
          const size_t rbase = (static_cast<size_t>(b) * heads + h) * Sq;
          const float2 m0 = drop_prob2(drop, rbase + i0, Sk, j), m1 = drop_prob2(drop, rbase + i1, Sk, j);
          dp[nt][0] *= m0.x; dp[nt][1] *= m0.y; dp[nt][2] *= m1.x; dp[nt][3] *= m1.y;
          constexpr int kD = DROP ? 1 : 0;
          pm[nt * kD][0] = p[nt][0] * m0.x; pm[nt * kD][1] = p[nt][1] * m0.y;
          pm[nt * kD][2] = p[nt][2] * m1.x; pm[nt * kD][3] = p[nt][3] * m1.y;
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          dot0 += p[nt][e] * dp[nt][e];
          dot1 += p[nt][2 + e] * dp[nt][2 + e];
        }
      }
    }
    dot0 = quad_sum(dot0); dot1 = quad_sum(dot1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * nk16) {
        dp[nt][0] = p[nt][0] * (dp[nt][0] - dot0) * 0.125f;     // now dS
        dp[nt][1] = p[nt][1] * (dp[nt][1] - dot0) * 0.125f;
        dp[nt][2] = p[nt][2] * (dp[nt][2] - dot1) * 0.125f;
        dp[nt][3] = p[nt][3] * (dp[nt][3] - dot1) * 0.125f;
      }
    }
    // dQ = dS·K
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < nk16) {
        uint32_t ah[4], al[4];
        split2(dp[2 * kk][0], dp[2 * kk][1], ah[0], al[0]);
        split2(dp[2 * kk][2], dp[2 * kk][3], ah[1], al[1]);
        split2(dp[2 * kk + 1][0], dp[2 * kk + 1][1], ah[2], al[2]);
        split2(dp[2 * kk + 1][2], dp[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bh[4], bl[4];
          load_b_kmajor(bh, Kh, kk * 16, np * 16, lane);
          load_b_kmajor(bl, Kl, kk * 16, np * 16, lane);
          mma3(dq[2 * np], ah, al, bh[0], bh[1], bl[0], bl[1]);
          mma3(dq[2 * np + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = h * 64 + nt * 8 + 2 * t;
      uint32_t hi, lo;
      if (i0 < Sq) {
        split2(dq[nt][0], dq[nt][1], hi, lo);
        *reinterpret_cast<uint32_t*>(dq_hi + (qrow0 + i0) * ld_d + c) = hi;
        *reinterpret_cast<uint32_t*>(dq_lo + (qrow0 + i0) * ld_d + c) = lo;
      }
      if (i1 < Sq) {
        split2(dq[nt][2], dq[nt][3], hi, lo);
        *reinterpret_cast<uint32_t*>(dq_hi + (qrow0 + i1) * ld_d + c) = hi;
        *reinterpret_cast<uint32_t*>(dq_lo + (qrow0 + i1) * ld_d + c) = lo;
      }
    }
  }
  __syncthreads();     // every warp is done with the V and K tiles: P and dS may overwrite them
  if (rows_live) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * nk16) {
        const int j = nt * 8 + 2 * t;
        uint32_t hi, lo;
        constexpr int kD = DROP ? 1 : 0;
        split2(DROP ? pm[nt * kD][0] : p[nt][0], DROP ? pm[nt * kD][1] : p[nt][1], hi, lo);
        *reinterpret_cast<uint32_t*>(Ph + swz(i0, j)) = hi; *reinterpret_cast<uint32_t*>(Pl + swz(i0, j)) = lo;
        split2(DROP ? pm[nt * kD][2] : p[nt][2], DROP ? pm[nt * kD][3] : p[nt][3], hi, lo);
        *reinterpret_cast<uint32_t*>(Ph + swz(i1, j)) = hi; *reinterpret_cast<uint32_t*>(Pl + swz(i1, j)) = lo;
        split2(dp[nt][0], dp[nt][1], hi, lo);
        *reinterpret_cast<uint32_t*>(Sh + swz(i0, j)) = hi; *reinterpret_cast<uint32_t*>(Sl + swz(i0, j)) = lo;
        split2(dp[nt][2], dp[nt][3], hi, lo);
        *reinterpret_cast<uint32_t*>(Sh + swz(i1, j)) = hi; *reinterpret_cast<uint32_t*>(Sl + swz(i1, j)) = lo;
      }
    }
  }
  __syncthreads();

  // phase 2 (warp = 16 key rows): dV = Pᵀ·dO, dK = dSᵀ·Q, contraction over the (zero-padded) query rows
  if (r0 < 16 * nk16) {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
      dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < nq16) {
        uint32_t ph[4], pl[4], sh[4], sl[4];
        load_a_trans(ph, Ph, r0, kk * 16, lane);
        load_a_trans(pl, Pl, r0, kk * 16, lane);
        load_a_trans(sh, Sh, r0, kk * 16, lane);
        load_a_trans(sl, Sl, r0, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bh[4], bl[4];
          load_b_kmajor(bh, Oh, kk * 16, np * 16, lane);
          load_b_kmajor(bl, Ol, kk * 16, np * 16, lane);
          mma3(dv[2 * np], ph, pl, bh[0], bh[1], bl[0], bl[1]);
          mma3(dv[2 * np + 1], ph, pl, bh[2], bh[3], bl[2], bl[3]);
          load_b_kmajor(bh, Qh, kk * 16, np * 16, lane);
          load_b_kmajor(bl, Ql, kk * 16, np * 16, lane);
          mma3(dk[2 * np], sh, sl, bh[0], bh[1], bl[0], bl[1]);
          mma3(dk[2 * np + 1], sh, sl, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = h * 64 + nt * 8 + 2 * t;
      uint32_t hi, lo;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = half ? i1 : i0;
        if (j < Sk) {
          const size_t idx = (krow0 + j) * ld_d + c;
          split2(dv[nt][2 * half], dv[nt][2 * half + 1], hi, lo);
          *reinterpret_cast<uint32_t*>(dv_hi + idx) = hi; *reinterpret_cast<uint32_t*>(dv_lo + idx) = lo;
          split2(dk[nt][2 * half], dk[nt][2 * half + 1], hi, lo);
          *reinterpret_cast<uint32_t*>(dk_hi + idx) = hi; *reinterpret_cast<uint32_t*>(dk_lo + idx) = lo;
        }
      }
    }
  }
}

}  // namespace

namespace {
int operand_maps(CUtensorMap* hi, CUtensorMap* lo, const AttnOperand& o) {
  int rc = tma_map_2d_bf16(hi, o.base.hi, o.ld, o.rows, o.ld, 64, 64, 128);
  if (rc) return rc;
  return tma_map_2d_bf16(lo, o.base.lo, o.ld, o.rows, o.ld, 64, 64, 128);
}
bool operand_ok(const AttnOperand& o) {
  return o.base.hi && o.base.lo && o.ld % 8 == 0 && o.col % 8 == 0 && o.rows > 0 &&
         !((reinterpret_cast<uintptr_t>(o.base.hi) | reinterpret_cast<uintptr_t>(o.base.lo)) & 15);
}
// one CTA per several samples: self-attention (Q, K, V rows of one matrix) with S ≤ 32
bool pack_self(const AttnOperand& q, const AttnOperand& k, const AttnOperand& v, int Sq, int Sk) {
  static const bool on = [] { const char* e = getenv("XLX_ATTN_PACK"); return !(e && e[0] == '0'); }();
  return on && Sq == Sk && Sq <= 32 && q.base.hi == k.base.hi && k.base.hi == v.base.hi && q.rows == k.rows &&
         k.rows == v.rows && q.ld == k.ld && k.ld == v.ld;
}
template <typename K>
int set_smem(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}
}  // namespace

int attention_fwd(AttnOperand q, AttnOperand k, AttnOperand v, const float* mask, int B, int heads, int Sq, int Sk,
                  Split ctx, float* ctx_f32, int ld_ctx, float* probs, cudaStream_t s, DropSite drop) {
  if (Sq < 1 || Sk < 1 || Sq > 64 || Sk > 64) return -5;
  if (!operand_ok(q) || !operand_ok(k) || !operand_ok(v) || (ld_ctx % 8)) return -2;
  if (!B) return 0;
  constexpr size_t smem = 6 * TT * sizeof(bf16) + 1024;
  static bool set = false;
  if (!set) {
    int rc = set_smem(attn_fwd_mma_kernel<false, false>, smem);
    if (!rc) rc = set_smem(attn_fwd_mma_kernel<true, false>, smem);
    if (!rc) rc = set_smem(attn_fwd_mma_kernel<false, true>, smem);
    if (!rc) rc = set_smem(attn_fwd_mma_kernel<true, true>, smem);
    if (rc) return rc;
    set = true;
  }
  FwdMaps maps;
  int rc;
  if ((rc = operand_maps(&maps.m[0], &maps.m[1], q))) return rc;
  if ((rc = operand_maps(&maps.m[2], &maps.m[3], k))) return rc;
  if ((rc = operand_maps(&maps.m[4], &maps.m[5], v))) return rc;
  const bool pack = pack_self(q, k, v, Sq, Sk);
  const int per = pack ? 64 / Sq : 1;
  const dim3 grid(heads, (B + per - 1) / per);
  auto kernel = pack ? (drop.threshold ? attn_fwd_mma_kernel<true, true> : attn_fwd_mma_kernel<false, true>)
                     : (drop.threshold ? attn_fwd_mma_kernel<true, false> : attn_fwd_mma_kernel<false, false>);
  launch_pdl(kernel, grid, dim3(AT), smem, s, maps, q.col, k.col, v.col, mask, heads, Sq, Sk, ctx.hi, ctx.lo, ctx_f32, ld_ctx,
             probs, drop, B);
  count_aux_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int attention_bwd(AttnOperand dctx, AttnOperand q, AttnOperand k, AttnOperand v, const float* probs, int B, int heads,
                  int Sq, int Sk, Split dq, Split dk, Split dv, int ld_d, cudaStream_t s, DropSite drop) {
  if (Sq < 1 || Sk < 1 || Sq > 64 || Sk > 64) return -5;
  if (!operand_ok(q) || !operand_ok(k) || !operand_ok(v) || !operand_ok(dctx) || (ld_d % 2)) return -2;
  if (!B) return 0;
  constexpr size_t smem = 8 * TT * sizeof(bf16) + 1024;
  static bool set = false;
  if (!set) {
    int rc = set_smem(attn_bwd_mma_kernel<false, false>, smem);
    if (!rc) rc = set_smem(attn_bwd_mma_kernel<true, false>, smem);
    if (!rc) rc = set_smem(attn_bwd_mma_kernel<false, true>, smem);
    if (!rc) rc = set_smem(attn_bwd_mma_kernel<true, true>, smem);
    if (rc) return rc;
    set = true;
  }
  BwdMaps maps;
  int rc;
  if ((rc = operand_maps(&maps.m[0], &maps.m[1], q))) return rc;
  if ((rc = operand_maps(&maps.m[2], &maps.m[3], k))) return rc;
  if ((rc = operand_maps(&maps.m[4], &maps.m[5], v))) return rc;
  if ((rc = operand_maps(&maps.m[6], &maps.m[7], dctx))) return rc;
  const bool pack = pack_self(q, k, v, Sq, Sk) && dctx.rows == q.rows;
  const int per = pack ? 64 / Sq : 1;
  const dim3 grid(heads, (B + per - 1) / per);
  auto kernel = pack ? (drop.threshold ? attn_bwd_mma_kernel<true, true> : attn_bwd_mma_kernel<false, true>)
                     : (drop.threshold ? attn_bwd_mma_kernel<true, false> : attn_bwd_mma_kernel<false, false>);
  launch_pdl(kernel, grid, dim3(AT), smem, s, maps, q.col, k.col, v.col, dctx.col, probs, heads, Sq, Sk, dq.hi, dq.lo, dk.hi,
             dk.lo, dv.hi, dv.lo, ld_d, drop, B);
  count_aux_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace xlx
