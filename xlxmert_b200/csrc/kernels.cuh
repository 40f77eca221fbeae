// Non-GEMM kernels of the X-LXMERT hot path (sm_100a): split-bf16 conversion, LayerNorm fwd/bwd,
// the per-(sample, head) attention core fwd/bwd, visual-feature embedding, column reductions.
// All launches are asynchronous on the given stream; none allocates.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xlx {

typedef __nv_bfloat16 bf16;

// A split-bf16 matrix: x ≈ hi + lo (see xlx_ptx.cuh).  lo may be null for "hi only" consumers.
struct Split {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

long long aux_launch_count();

// x[n] (fp32) → hi/lo.  n must be a multiple of 4.
int split_f32(const float* x, Split out, size_t n, cudaStream_t s);
// rows gathered from a table: out[r, :] = (mask && mask[r]) ? fill[:] : table[ids[r], :]  → f32 (opt) + split.
int gather_rows_split(const float* table, const int64_t* ids, const uint8_t* mask, const float* fill, int rows,
                      int cols, float* out_f32, Split out, cudaStream_t s);
// concat 3 fp32 vectors of length n each into dst[3n]
int concat3_f32(const float* a, const float* b, const float* c, float* dst, int n, cudaStream_t s);
// hi + lo → fp32
int unsplit_f32(Split in, float* out, size_t n, cudaStream_t s);

// LayerNorm over the last axis (biased variance), HF modeling_lxmert.py:188,281,343 (eps 1e-12).
//   y [M,H] fp32 → split out (+ optional fp32 out); optionally saves mean / rstd per row.
int layernorm_fwd(const float* y, const float* gamma, const float* beta, float eps, int M, int H, Split out,
                  float* out_f32, float* mean, float* rstd, cudaStream_t s);
// dy [M,H] (upstream grad wrt LN output), y (saved LN input), mean/rstd → dx fp32 (+ split), and partial
// column sums for dgamma / dbeta: part[2, nblk, H]; call colsum_finish afterwards.
int layernorm_bwd(const float* dy, const float* y, const float* gamma, const float* mean, const float* rstd, int M,
                  int H, float* dx, Split dx_split, float* part, int* nblk_out, cudaStream_t s);
// reduce part[nvec, nblk, H] over nblk → out_a[H] (vec 0), out_b[H] (vec 1, optional)
int colsum_finish(const float* part, int nblk, int H, float* out_a, float* out_b, cudaStream_t s);
int layernorm_bwd_max_blocks(int M);

// column sums of a split (or fp32) matrix: out[N] = Σ_m x[m, n]  (bias gradients).  scratch: [colsum_blocks(M), N].
int colsum(const float* x_f32, Split x, int M, int N, int ld, float* scratch, float* out, cudaStream_t s);
int colsum_blocks(int M);

// LxmertVisualFeatureEncoder tail (HF:476-484): out = (LN_v(y1) + LN_b(pos·Wpᵀ + bp)) / 2.
// y1 = visn_fc(feats) (fp32, bias included) [M,H]; pos [M,4].
int visn_embed_fwd(const float* y1, const float* pos, const float* Wp, const float* bp, const float* g1,
                   const float* b1, const float* g2, const float* b2, float eps, int M, int H, Split out,
                   float* out_f32, float* stats /* [4,M]: mean1,rstd1,mean2,rstd2 or null */, cudaStream_t s);
// backward: dout [M,H] → dy1 (fp32 + split), partial sums part[7, nblk, H]: dg1, db1, dg2, db2, dbp, and
// dWp as 4 vectors ([H] each for the 4 box coordinates) → part has 9 vectors total.
int visn_embed_bwd(const float* dout, const float* y1, const float* pos, const float* Wp, const float* bp,
                   const float* g1, const float* g2, const float* stats, int M, int H, float* dy1, Split dy1_split,
                   float* part, int* nblk_out, cudaStream_t s);
int colsum_finish_n(const float* part, int nvec, int nblk, int H, float* const* outs, cudaStream_t s);

// Attention core, LxmertAttention.forward (HF:238-274) after the projections, head_dim 64:
//   P = softmax(Q·Kᵀ/8 + mask), ctx = P·V per (sample, head).
// q/k/v: fp32 matrices with row stride ld (elements); sample b's query rows start at row b*Sq, key/value rows
// at b*Sk; head h occupies columns [h*64, h*64+64).  mask: additive fp32 [B, Sk] or null.
// ctx written as split rows [b*Sq + i, h*64 + d] with leading dimension ld_ctx.  probs (optional) [B,heads,Sq,Sk].
int attention_fwd(const float* q, const float* k, const float* v, int ld, const float* mask, int B, int heads,
                  int Sq, int Sk, Split ctx, float* ctx_f32, int ld_ctx, float* probs, cudaStream_t s);
// Backward: dctx fp32 [B*Sq, ld_dctx]; writes dq (rows b*Sq+i), dk/dv (rows b*Sk+j) as split (and fp32 if given)
// with leading dimension ld_d.
int attention_bwd(const float* dctx, int ld_dctx, const float* q, const float* k, const float* v, int ld,
                  const float* probs, int B, int heads, int Sq, int Sk, Split dq, Split dk, Split dv, int ld_d,
                  cudaStream_t s);

}  // namespace xlx
