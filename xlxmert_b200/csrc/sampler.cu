// Device side of the text→image sampling loops: the index / select work the
// reference does with torch.topk + scatter_ + torch.where + nn.Embedding between two encoder passes
// (x-lxmert/src/tasks/imggen_model.py:204-218,238-243 for mask-predict, :135-153 for the one-cell-per-step loop).
// One kernel per transition; nothing returns to the host, so a whole sampling run is capturable in one CUDA graph.
// Integer / byte work + row copies: bit-exact by construction, HBM-bound (B·64 rows of 8 KB).
//
// Tie rule: torch.topk's order among EQUAL probabilities is implementation-defined (CPU and CUDA differ); here the lower
// cell index ranks first, deterministically.
#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

using namespace xlx;

namespace {

constexpr int kMaxCells = 64;

// block = one sample; 256 threads copy rows with float4 accesses
__device__ __forceinline__ void copy_row(float* dst, const float* src, int F) {
  for (int c = threadIdx.x * 4; c < F; c += blockDim.x * 4)
    *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(src + c));
}

// Mask-predict transition.  pred_prob == nullptr: initial state (imggen_model.py:204-206,215-218 at i = 0): every cell
// masked, code = mask_feat.  Otherwise:
//   code[b,i] ← vis_mask[b,i] ? table[pred_id[b,i]] : code[b,i]                      (:238-243)
//   next mask = the n_next cells of LOWEST pred_prob (topk(largest=False) + scatter_)  (:209-212)
//   code[b,i] ← next_mask[b,i] ? mask_feat : code[b,i]                                (:215-218 of the next iteration)
__global__ void __launch_bounds__(256)
nar_update_kernel(float* code, const uint8_t* vis_mask, const float* pred_prob, const int64_t* pred_id,
                  const float* table, const float* mask_feat, int V, int F, int n_next, uint8_t* vis_mask_next) {
  __shared__ float s_p[kMaxCells];
  __shared__ uint8_t s_next[kMaxCells];
  const int b = blockIdx.x;
  if (pred_prob) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) s_p[i] = pred_prob[static_cast<size_t>(b) * V + i];
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float p = s_p[i];
      int rank = 0;
      for (int j = 0; j < V; ++j) rank += (s_p[j] < p) || (s_p[j] == p && j < i);
      s_next[i] = rank < n_next ? 1 : 0;
    }
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) s_next[i] = 1;
  }
  __syncthreads();
  for (int i = 0; i < V; ++i) {
    const size_t r = static_cast<size_t>(b) * V + i;
    float* dst = code + r * F;
    if (s_next[i]) copy_row(dst, mask_feat, F);
    else if (pred_prob && vis_mask[r]) copy_row(dst, table + static_cast<size_t>(pred_id[r]) * F, F);
  }
  __syncthreads();     // vis_mask_next may alias vis_mask: every read above is done before any write below
  for (int i = threadIdx.x; i < V; i += blockDim.x) vis_mask_next[static_cast<size_t>(b) * V + i] = s_next[i];
}

// One-cell-per-step transition (imggen_model.py:135-153).  position ≥ 0: that cell (raster / random order, :137-139);
// position < 0: the unvisited cell of highest pred_prob (masked_fill(visited, −10000) → topk(1), :140-149).
//   code[b,top] ← table[pred_id[b,top]];  vis_mask[b,top] ← 0;  visited[b,top] ← 1 (confidence order only)
__global__ void __launch_bounds__(256)
ar_update_kernel(float* code, uint8_t* vis_mask, uint8_t* visited, const float* pred_prob, const int64_t* pred_id,
                 const float* table, int V, int F, int position) {
  __shared__ int s_top;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int top = position;
    if (position < 0) {
      float best = -INFINITY;
      top = 0;
      for (int i = 0; i < V; ++i) {
        const size_t r = static_cast<size_t>(b) * V + i;
        const float p = visited[r] ? -10000.0f : pred_prob[r];
        if (p > best) { best = p; top = i; }
      }
    }
    s_top = top;
  }
  __syncthreads();
  const int top = s_top;
  const size_t r = static_cast<size_t>(b) * V + top;
  copy_row(code + r * F, table + static_cast<size_t>(pred_id[r]) * F, F);
  if (threadIdx.x == 0) {
    vis_mask[r] = 0;
    if (position < 0) visited[r] = 1;
  }
}

// code[b, cell] ← mask_feat, vis_mask[b, cell] ← 1 for every sample (imggen_model.py:111-113: a revisited position is
// masked again before the pass)
__global__ void __launch_bounds__(256)
remask_cell_kernel(float* code, uint8_t* vis_mask, const float* mask_feat, int V, int F, int cell) {
  const size_t r = static_cast<size_t>(blockIdx.x) * V + cell;
  copy_row(code + r * F, mask_feat, F);
  if (threadIdx.x == 0) vis_mask[r] = 1;
}

int check(int B, int V, int F) {
  if (B < 1 || V < 1 || F < 4) return -21;
  if (V > kMaxCells) return -22;
  if (F % 4) return -2;
  return 0;
}

}  // namespace

extern "C" {

int32_t xlx_sampler_nar_update(float* code, const uint8_t* vis_mask, const float* pred_prob, const int64_t* pred_id,
                               const float* table, const float* mask_feat, int32_t B, int32_t V, int32_t F,
                               int32_t n_mask_next, uint8_t* vis_mask_next, void* stream) {
  XLX_TRY(check(B, V, F));
  if (!code || !mask_feat || !vis_mask_next) return -24;
  if (pred_prob && (!vis_mask || !pred_id || !table)) return -24;
  if (n_mask_next < 0 || n_mask_next > V) return -1;
  if ((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(mask_feat)) & 15)
    return -2;
  XLX_TRY(ensure_device(code));
  nar_update_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(code, vis_mask, pred_prob, pred_id, table,
                                                                      mask_feat, V, F, n_mask_next, vis_mask_next);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}

int32_t xlx_sampler_ar_update(float* code, uint8_t* vis_mask, uint8_t* visited, const float* pred_prob,
                              const int64_t* pred_id, const float* table, int32_t B, int32_t V, int32_t F,
                              int32_t position, void* stream) {
  XLX_TRY(check(B, V, F));
  if (!code || !vis_mask || !pred_id || !table) return -24;
  if (position >= V) return -1;
  if (position < 0 && (!visited || !pred_prob)) return -24;
  if ((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(table)) & 15) return -2;
  XLX_TRY(ensure_device(code));
  ar_update_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(code, vis_mask, visited, pred_prob, pred_id, table,
                                                                     V, F, position);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}

int32_t xlx_sampler_remask_cell(float* code, uint8_t* vis_mask, const float* mask_feat, int32_t B, int32_t V, int32_t F,
                                int32_t cell, void* stream) {
  XLX_TRY(check(B, V, F));
  if (!code || !vis_mask || !mask_feat) return -24;
  if (cell < 0 || cell >= V) return -1;
  if ((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(mask_feat)) & 15) return -2;
  XLX_TRY(ensure_device(code));
  remask_cell_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(code, vis_mask, mask_feat, V, F, cell);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
