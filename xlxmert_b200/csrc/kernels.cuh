// Non-GEMM kernels of the X-LXMERT hot path (sm_100a): split-bf16 conversion, LayerNorm fwd/bwd,
// the per-(sample, head) attention core fwd/bwd, box-position linear, column reductions.
// All launches are asynchronous on the given stream; none allocates.  Return 0 or a CUDA error code
// (positive) / argument error (negative).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dropout.cuh"
#include "pdl.cuh"

namespace xlx {

typedef __nv_bfloat16 bf16;

// A split-bf16 matrix: x ≈ hi + lo (see xlx_ptx.cuh).
struct Split {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

long long aux_launch_count();

// x[n] (fp32) → hi/lo.  n must be a multiple of 4.
int split_f32(const float* x, Split out, size_t n, cudaStream_t s);
// hi + lo → fp32
int unsplit_f32(Split in, float* out, size_t n, cudaStream_t s);
// out[r, :] = (mask && mask[r]) ? fill[:] : table[ids[r], :]  → fp32 (optional) + split (optional).
// The X-LXMERT visual input: vis_emb(cluster_ids) then torch.where(vis_mask, mask_feat, feats)
// (lxrt/modeling.py:185-193).  cols % 4 == 0.
int gather_rows(const float* table, const int64_t* ids, const uint8_t* mask, const float* fill, int rows, int cols,
                float* out_f32, Split out, cudaStream_t s);
// dst[3n] = a[n] | b[n] | c[n]
int concat3_f32(const float* a, const float* b, const float* c, float* dst, int n, cudaStream_t s);
int fill_f32(float* dst, float v, size_t n, cudaStream_t s);

// Batched weight preparation: for every job, src [R, C] fp32 → dst rows [row0 .., row0 + R) of a split matrix with
// leading dimension C (dst may be null), and → its transpose dst_t [C, ld_t] at column offset col0_t (dst_t may be
// null); up to 192 jobs per launch, the job table travels as a kernel parameter (no device-side table, no copies).
struct SplitJob {
  const float* src;
  bf16 *hi, *lo;        // row-major copy: element (r, c) at (row0 + r) * C + c
  bf16 *t_hi, *t_lo;    // transposed copy: element (r, c) at c * ld_t + col0_t + r
  int R, C, row0, ld_t, col0_t, tile0;   // tile0: first 64×64 tile index of this job (prefix sum, filled by the callee)
};
int split_batch(SplitJob* jobs, int njobs, cudaStream_t s);

// x rows with stride ld_in (fp32) → contiguous split matrix [rows, cols]; cols % 4 == 0.
int split_rows_f32(const float* x, size_t ld_in, int rows, int cols, Split out, cudaStream_t s);
// dpre = d · (1 − y²) (tanh backward) → split; n % 4 == 0.
int tanh_bwd_split(const float* d, const float* y, Split out, size_t n, cudaStream_t s);

// LxmertEmbeddings (HF modeling_lxmert.py:191-214) before its LayerNorm:
//   y[r,:] = word[ids[r],:] + pos[r % L,:] + type[tt ? tt[r] : 0,:]      r < B·L, H % 4 == 0
int embed_sum(const int64_t* ids, const int64_t* tt, const float* word, const float* pos, const float* type, int rows,
              int L, int H, float* y, cudaStream_t s);
// Scatter-add of dy [rows,H] into the three (pre-zeroed) table gradients; row 0 of each table has
// padding_idx semantics (no gradient, HF:184-186).
int embed_scatter(const int64_t* ids, const int64_t* tt, const float* dy, int rows, int L, int H, float* dword,
                  float* dpos, float* dtype, cudaStream_t s);

// du = dg ∘ gelu'(u) → split; n % 4 == 0.
int gelu_bwd_split(const float* dg, const float* u, Split out, size_t n, cudaStream_t s);

// ---- cross-entropy / arg-max over a [M, C] logits matrix with row stride ld (fp32) ---------------------
// torch.nn.CrossEntropyLoss() (mean over rows whose label != ignore, lxrt/modeling.py:99,102,253-256):
// lse[m] = logsumexp(logits[m,:]); stats[0] = loss, stats[1] = number of valid rows (as float).
int ce_fwd(const float* logits, int ld, int M, int C, const int64_t* labels, int64_t ignore_index, float* lse,
           float* rowloss, float* stats, cudaStream_t s);
// dlogits[m,c] = d_loss · (softmax(logits)[m,c] − 1[c == label]) / n_valid for valid rows, else 0; columns
// C..Cp-1 are written as zero.  Output: split matrix [M, Cp].
int ce_bwd(const float* logits, int ld, int M, int C, int Cp, const int64_t* labels, int64_t ignore_index,
           const float* lse, const float* stats, const float* d_loss, Split dlogits, cudaStream_t s);
// Feature-regression loss of the cluster head (lxrt/modeling.py:270-284): SmoothL1(β = 1) per element, mean over the
// feature axis, masked mean per sample, batch mean — folded into one weight per row (feat_row_weight):
//   loss = Σ_r w[r] · Σ_f SmoothL1(feat[r,f] − target[r,f]);   rowloss [M] scratch, loss = device scalar.
int smooth_l1_fwd(const float* feat, const float* target, const float* w, int M, int F, float* rowloss, float* loss,
                  cudaStream_t s);
// out[r,f] = d_loss · w[r] · clamp(feat − target, −1, 1); out may alias feat.
int smooth_l1_bwd(const float* feat, const float* target, const float* w, const float* d_loss, int M, int F, float* out,
                  cudaStream_t s);
// w[i] = vis_mask[row_i] / (B · max(n_mask[b], 1) · F), row_i = rows ? rows[i] : i  (rows: compacted row indices).
int feat_row_weight(const uint8_t* vis_mask, int B, int V, int F, const int64_t* rows, int n, float* w, cudaStream_t s);
// fp32 [M, C] → split [M, Cp] with zero padding columns.
int split_pad_f32(const float* x, int M, int C, int Cp, Split out, cudaStream_t s);
// softmax(logits, -1).max(-1) → (prob, id); ties go to the first index like torch.max
// (tasks/imggen_model.py:232-235).
int softmax_argmax(const float* logits, int ld, int M, int C, float* prob, int64_t* id, cudaStream_t s);

// Finish of a GEMM row-statistics epilogue (gemm_sm100.cuh: GemmEpilogue::rowstat): rowstat [M, slots, 3] partial
// (max, Σ exp(x − max), first index of the max) → prob[m] = softmax(row).max() = 1 / Σ_j exp(x_j − max), id[m] = its index.
// Index rule = softmax_argmax's (torch.max over the PROBABILITIES, first index): among the partial maxima whose
// probability rounds to the row maximum's (expf(m_p − max) == 1.0f) the lowest index wins.  Identical to the
// materialised kernel unless two DISTINCT logits of a row lie within 2^-25 of its maximum inside one partial — which
// fp32 spacing rules out whenever max|logit| ≥ 1.
int rowstat_merge(const float* rowstat, int M, int slots, float* prob, int64_t* id, cudaStream_t s);

// Same loss for a tiny class count (C ≤ 8, row stride C): rowloss/stats as ce_fwd; dlogits fp32 [M,C].
int small_ce_fwd(const float* logits, int M, int C, const int64_t* labels, int64_t ignore_index, float* rowloss,
                 float* stats, cudaStream_t s);
int small_ce_bwd(const float* logits, int M, int C, const int64_t* labels, int64_t ignore_index, const float* stats,
                 const float* d_loss, float* dlogits, cudaStream_t s);

// y[M,N] = x[M,K]·W[N,K]ᵀ + b for tiny N (≤ 8) — the 2-way matched head (HF:661,664).
int small_linear_fwd(const float* x, const float* W, const float* b, int M, int K, int N, float* y, cudaStream_t s);
// dW[N,K] = dyᵀ·x, db[N] = Σ dy, dx[M,K] = dy·W.
int small_linear_bwd(const float* dy, const float* x, const float* W, int M, int K, int N, float* dW, float* db,
                     float* dx, cudaStream_t s);

// LayerNorm over the last axis (biased variance; HF modeling_lxmert.py:188,281,343 use eps 1e-12):
//   o = out_scale · LN(y) + addend;   y [M,H] fp32 → split out and/or fp32 out; optionally saves mean / rstd.
//   `drop` (optional): dropout applied to o (after the addend) — LxmertEmbeddings / LxmertVisualFeatureEncoder outputs.
int layernorm_fwd(const float* y, const float* gamma, const float* beta, float eps, int M, int H, float out_scale,
                  const float* addend, Split out, float* out_f32, float* mean, float* rstd, cudaStream_t s,
                  DropSite drop = DropSite());
// dy [M,H]: upstream grad wrt the LN output (scaled by dy_scale); y: saved LN input; → dx fp32 and/or split,
// plus partial column sums part[3, nblk, H] (dgamma, dbeta, Σ_rows dx); finish with colsum_finish.  dx may alias dy.
// Dropout (training): `drop_in` gates the incoming dy with the mask the forward applied to this LayerNorm's output;
// `drop_out` is the mask of a dropout(dense(x)) + residual producer of the LayerNorm INPUT: dx / dx_split keep the plain
// gradient (residual path), dx_masked receives dx ∘ mask (the dense path's GEMM operand) and the third partial column
// sum becomes Σ_rows dx ∘ mask (the dense bias gradient).
int layernorm_bwd(const float* dy, float dy_scale, const float* y, const float* gamma, const float* mean,
                  const float* rstd, int M, int H, float* dx, Split dx_split, float* part, int* nblk_out,
                  cudaStream_t s, DropSite drop_in = DropSite(), DropSite drop_out = DropSite(),
                  Split dx_masked = Split());
// out[rows, H] / out[rows, Sk] = the multipliers (0 or 1/(1−p)) the fused kernels apply at a site (tests).
int dropout_mask_hidden(DropSite d, size_t rows, int H, float* out, cudaStream_t s);
int dropout_mask_probs(DropSite d, size_t rows, int Sk, float* out, cudaStream_t s);
int reduce_max_blocks();  // upper bound on nblk for every partial-sum kernel here
// out_v[H] = Σ_blk part[v, blk, H] for v < nvec (outs[v] may be null to skip); accumulate: out += instead of =
int colsum_finish(const float* part, int nvec, int nblk, int H, float* const* outs, int accumulate, cudaStream_t s);

// Deferred, batched finish of partial column sums: a backward pass produces one small reduction per block (LayerNorm
// affine + dense-bias gradients, intermediate-bias and Q/K/V-bias gradients) — ≈ 140 nine-microsecond launches per
// step when finished one by one.  The partials stay where their producers wrote them (each in its own piece of a
// scratch arena) and ONE launch per 64 jobs finishes them all at the end of the pass, in the same fixed order.
struct FinishJob {
  const float* part;   // [nvec, nblk, H]
  float* out[3];       // out[v][h] = Σ_blk part[v, blk, h]   (nullptr to skip a vector)
  int nvec, nblk, H;
  int block0;          // first CTA of this job (prefix sum, filled by colsum_finish_batched)
};
int colsum_finish_batched(FinishJob* jobs, int njobs, cudaStream_t s);
// first phase of colsum(): partial sums [nblk, N] into `part`; *nblk_out = rows of partials written
int colsum_partial(const float* x_f32, Split x, int M, int N, int ld, float* part, int* nblk_out, cudaStream_t s,
                   const uint8_t* rowmask = nullptr);

// column sums: out[N] = Σ_m x[m, n]  (bias gradients) of an fp32 or split matrix with leading dimension ld.
// scratch: [128, N] floats.  rowmask (optional, one byte per row): only rows with a non-zero byte are summed.
int colsum(const float* x_f32, Split x, int M, int N, int ld, float* scratch, float* out, cudaStream_t s,
           const uint8_t* rowmask = nullptr);

// box branch of LxmertVisualFeatureEncoder (HF:479-480): y2[m,h] = bp[h] + Σ_j pos[m,j]·Wp[h,j], j < 4.
int box_linear_fwd(const float* pos, const float* Wp, const float* bp, int M, int H, float* y2, cudaStream_t s);
// dWp[h,j] = Σ_m dy2[m,h]·pos[m,j], dbp[h] = Σ_m dy2[m,h]; scratch [5, reduce_max_blocks(), H].
int box_linear_bwd(const float* dy2, const float* pos, int M, int H, float* scratch, float* dWp, float* dbp,
                   cudaStream_t s);

// Attention core, LxmertAttention.forward (HF:238-274) after the projections, head_dim 64 (attention.cu):
//   P = softmax(Q·Kᵀ/8 + mask), ctx = P·V per (sample, head).
// An operand is a split-bf16 matrix [rows, ld]; sample b uses rows b·S.. and head h the columns col + h·64 ..
struct AttnOperand {
  Split base;
  int ld = 0, rows = 0, col = 0;
};
// mask: additive fp32 [B, Sk] or null.  Sq, Sk ≤ 64.  ctx: split and/or fp32 rows [b·Sq + i, h·64 + d] with leading
// dimension ld_ctx.  probs (optional) [B, heads, Sq, Sk] fp32.
// `drop` (training): dropout on the probabilities before P·V (HF:262-266); `probs` always receives the UNdropped P.
int attention_fwd(AttnOperand q, AttnOperand k, AttnOperand v, const float* mask, int B, int heads, int Sq, int Sk,
                  Split ctx, float* ctx_f32, int ld_ctx, float* probs, cudaStream_t s, DropSite drop = DropSite());
// Backward: dctx = gradient wrt ctx (an operand like q); writes dq (rows b·Sq + i), dk / dv (rows b·Sk + j), columns
// h·64 + d, as split matrices with leading dimension ld_d.
int attention_bwd(AttnOperand dctx, AttnOperand q, AttnOperand k, AttnOperand v, const float* probs, int B, int heads,
                  int Sq, int Sk, Split dq, Split dk, Split dv, int ld_d, cudaStream_t s, DropSite drop = DropSite());
// Row compaction for the masked-prediction losses (CrossEntropyLoss ignores label −100, lxrt/modeling.py:99,253-256):
// rows[0..count) = ascending indices m with labels[m] != ignore.  rows has room for M entries.
int labelled_rows(const int64_t* labels, int M, int64_t ignore, int64_t* rows, int32_t* count, cudaStream_t s);
// dst [M, cols] = 0, then dst[rows[i], :] = src[i, :] for i < n.
int scatter_rows(const float* src, const int64_t* rows, int n, int M, int cols, float* dst, cudaStream_t s);
int gather_i64(const int64_t* src, const int64_t* rows, int n, int64_t* dst, cudaStream_t s);
void count_aux_launch();


}  // namespace xlx
