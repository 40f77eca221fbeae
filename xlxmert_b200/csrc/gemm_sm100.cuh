// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] = epilogue( Σ_passes A_p[M,K] · B_p[N,K]ᵀ )
//
// Operands are "split-bf16" matrices (hi/lo bf16 pairs, x ≈ hi + lo): with passes = 3 the kernel
// issues A_hi·B_hi + A_lo·B_hi + A_hi·B_lo into one fp32 TMEM accumulator, which reproduces an fp32
// GEMM to ≈ 2^-16 relative (needed for the reference's fp32 parity bar);
// passes = 1 uses the hi parts only (plain bf16 tensor-core GEMM).
//
// Roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + UMMA issuer, warps 2-17 = epilogue
// (TMEM → registers → 64-byte-swizzled shared-memory transpose → coalesced global traffic or a TMA store).  Shared
// memory is a ring of stages filled by TMA (cp.async.bulk.tensor, hardware swizzle) and drained by tcgen05.mma; the
// accumulator is double buffered in TMEM (2 × BN fp32 columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Tiles are 128 × BN (BN ≤ 256) × 32; CTA pairs share the B tile by TMA multicast; weight-gradient
// shapes split K with a deterministic partial-sum reduce; kernels are launched with programmatic dependent launch.
// Convolutions run as implicit GEMMs over NHWC activations (ConvGeometry): one 4-D TMA box per filter tap, or — for
// 3×3 at width ≥ 128 — one haloed row per filter row whose three dx taps are shifted descriptor views ("row halo").
//
// This is the workhorse behind every nn.Linear on the hot path (HF modeling_lxmert.py:217-350 Q/K/V,
// attention output, intermediate, output; lxrt/modeling.py:38-53 cluster head) and their dgrad / wgrad.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dropout.cuh"

namespace xlx {

enum EpiFlags : int {
  EPI_GELU = 1,        // v = gelu_erf(v) after bias (pre-activation optionally saved to out_u)
  EPI_GELU_GRAD = 2,   // v *= gelu'(u_in[m,n])
  EPI_ACCUM = 4,       // out_f32 += v instead of = v
  EPI_TANH = 8,        // v = tanh(v) after bias
  EPI_SAVE_DGELU = 16, // with EPI_GELU: out_u receives gelu'(pre-activation) instead of the pre-activation
  EPI_MUL = 32,        // v *= u_in[m,n]  (backward of an activation whose derivative was saved by the forward)
  EPI_RELU = 64,       // v = max(v, 0) after bias
};

struct GemmEpilogue {
  const float* bias = nullptr;      // [N]
  const float* addend = nullptr;    // [M, ld_addend] fp32, added last
  const float* u_in = nullptr;      // [M, ld_u] pre-activation for EPI_GELU_GRAD
  const __nv_bfloat16* addend_hi = nullptr;  // optional split addend (hi + lo), same ld as ld_addend
  const __nv_bfloat16* addend_lo = nullptr;
  float* out_f32 = nullptr;         // [M, ld_out]
  float* out_u = nullptr;           // [M, ld_u] pre-activation save (EPI_GELU)
  __nv_bfloat16* out_u16 = nullptr; // same save as bf16 (e.g. gelu' for the backward: a multiplier, 8 bits suffice)
  const __nv_bfloat16* u_in16 = nullptr;  // bf16 multiplier for EPI_MUL (instead of u_in)
  __nv_bfloat16* out_hi = nullptr;  // [M, ld_split]
  __nv_bfloat16* out_lo = nullptr;
  // Optional column sums of the final values (a bias gradient for free): every epilogue warp writes the sums over its
  // 32 rows, colsum_part[(m / 32), n] with m / 32 < 4·ceil(M/128); reduce the rows with colsum_finish (fixed order).
  float* colsum_part = nullptr;
  // Optional per-row soft-max statistics INSTEAD of any output (the sampler's softmax(logits).max(-1),
  // tasks/imggen_model.py:232-235, without ever writing the [M, classes] logits): every epilogue warp folds the columns
  // it reads of a tile (after alpha and bias) into (max, Σ exp(x − max), first index of the max) and writes them to
  // rowstat[(row · slots + slot) · 3 + {0,1,2}] (index stored as int bits), slot = n_tile · 4 + warp-of-quadrant,
  // slots = gemm_rowstat_slots(N).  Columns ≥ rowstat_cols (padding) are ignored.  Finish with rowstat_merge().
  float* rowstat = nullptr;
  int rowstat_cols = 0;
  // SPADE modulation epilogue (image_generator/src/layers.py:33-47,93-107) for the γ/β convolution (N = 64: columns
  // 0-31 = γ, 32-63 = β of the same 32 channels): instead of writing γ|β, the epilogue reads the block input x and writes
  //   out[p, c] = LeakyReLU_0.2( ((x[p, c] − mean[b, c])·rstd[b, c])·(1 + γ[p, c]) + β[p, c] + noise_w·noise[p] )
  // as fp32 (out_f32, ld_out) and/or split bf16 (out_hi/out_lo, ld_split) — γ|β never reach memory.  b = p >> hw_log2.
  const float* spade_x = nullptr;       // [pixels, spade_ldx] fp32, first 32 channels
  const float* spade_mean = nullptr;    // [B, 32] InstanceNorm statistics
  const float* spade_rstd = nullptr;
  const float* spade_noise = nullptr;   // [pixels] or null
  const float* spade_noise_w = nullptr; // device scalar (NoiseInjection.weight) or null
  int spade_ldx = 0, spade_hw_log2 = 0;
  // Optional dropout of the value after bias / activation and BEFORE the addends (LxmertAttentionOutput / LxmertOutput:
  // dropout(dense(x)) + residual, HF:282-287,344-349).  The GEMM must span the whole [M, N] matrix the site covers
  // (element index = row·N + col).  drop.threshold == 0: off.
  DropSite drop;
  int ld_addend = 0, ld_u = 0, ld_out = 0, ld_split = 0;
  int flags = 0;
  float alpha = 1.0f;               // v = alpha * acc before bias
};

struct GemmOperand {
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;   // may be null when passes == 1
  int ld = 0;                          // leading dimension in elements
  int mn_major = 0;                    // 0: stored [rows(M or N), K]; 1: stored [K, rows] (rows contiguous)
  int rows = 0;                        // rows that exist in memory (0 → M or N); tiles beyond are zero-filled by TMA
  int kext = 0;                        // extent along K that exists in memory (0 → K); zero-filled beyond
};

// Implicit-GEMM convolution: A is not a matrix but an NHWC activation tensor [B, H, W, C] (split bf16, a.hi/a.lo);
// GEMM row m = flat pixel index (b·H + y)·W + x, GEMM column k = tap·C + c with tap = ky·3 + kx (taps = 9, zero
// padding 1) or the pixel itself (taps = 1).  M = B·H·W, K = taps·C.  Each k-block is one TMA box of the tensor
// shifted by the tap offset, out-of-bounds zero-filled by the TMA unit — no im2col buffer exists anywhere.
// Constraints: C % 32 == 0, and H·W either divides 128 or is a multiple of 128 with W dividing 128 or a multiple.
struct ConvGeometry {
  int enabled = 0;
  int H = 0, W = 0, C = 0, taps = 9;
};

struct GemmProblem {
  int M = 0, N = 0, K = 0;
  GemmOperand a, b;
  ConvGeometry conv;
  int passes = 3;
  GemmEpilogue epi;
  // Optional split-K scratch (fp32).  When given, GEMMs with a plain fp32 epilogue whose output has too few
  // tiles to occupy the SMs (weight gradients: small M·N, huge K) are split along K into partial sums here
  // and reduced by a second kernel.  gemm_splitk_ws_floats() is always enough.
  float* splitk_ws = nullptr;
  size_t splitk_ws_floats = 0;
};
size_t gemm_splitk_ws_floats();
// The last 1024 words of a split-K workspace are per-tile arrival counters: the CTA that stores the LAST partial sum of
// an output tile adds the tile's partials up (fixed order: deterministic) and writes the result — no second kernel.
// The counters return to zero by themselves; a workspace must be reset once before its first use (and after an
// aborted launch): encoder / head backward calls do it once per call.
int gemm_splitk_ws_reset(float* splitk_ws, cudaStream_t stream);
// number of partial (max, sumexp, argmax) triples per row a rowstat epilogue writes for an N-column GEMM
int gemm_rowstat_slots(int N);
// Cached 2-D TMA descriptor of a row-major bf16 matrix [outer, inner] with leading dimension ld (elements) and a
// box of box_inner × box_outer elements; swizzle_bytes ∈ {0, 32, 64, 128}.  0 on success.
int tma_map_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer, int swizzle_bytes);

// Returns 0 on success, negative on invalid arguments, positive CUDA error otherwise.
int gemm_launch(const GemmProblem& p, cudaStream_t stream);
// Number of kernels launched by gemm_launch since process start (bench bookkeeping).
long long gemm_launch_count();
// Per-launch CUDA-event timing of every GEMM between begin and end (algorithmic FLOPs = 2·M·N·K each).
void gemm_timing_begin();
bool gemm_timing_active();   // per-launch event timing is on: callers keep everything on one stream
int gemm_timing_end(double* total_ms, double* total_flops, long long* launches);

}  // namespace xlx
