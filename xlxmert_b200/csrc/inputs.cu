// Input path of the pre-training step: the device half of Trainer.forward
// (x-lxmert/src/pretrain/lxmert_pretrain.py:143-225).  The reference issues up to 13 separate `.to(device)` copies
// per step and builds labels / masks with a dozen tiny torch kernels (`obj_labels[~vis_mask] = -100`, `word_id > 0`,
// `zeros_like`, …).  Here the host packs the step's arrays — exactly the ones collate_fn produces
// (lxmert_data.py:497-652) — into ONE pinned buffer, one cudaMemcpyAsync moves it, and one kernel unpacks it into the
// tensors XLxmertForPretraining.forward takes.  Integer / byte work: bit-exact by construction, HBM-bound
// (≈ 1.6 MB per 256-sample step), nothing for the tensor cores.
#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

using namespace xlx;

namespace {

constexpr int64_t kIgnore = -100;   // CrossEntropyLoss ignore_index (lxrt/modeling.py:99; lxmert_pretrain.py:163-166)

struct Layout {
  size_t word_id, word_label, matched_label, cluster_id, vis_mask, box_position, qa_label, bytes;
};
inline size_t up256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
Layout layout(int B, int L, int V) {
  Layout o;
  size_t off = 0;
  o.word_id = off;        off = up256(off + static_cast<size_t>(B) * L * 8);
  o.word_label = off;     off = up256(off + static_cast<size_t>(B) * L * 8);
  o.matched_label = off;  off = up256(off + static_cast<size_t>(B) * 8);
  o.cluster_id = off;     off = up256(off + static_cast<size_t>(B) * V * 8);
  o.vis_mask = off;       off = up256(off + static_cast<size_t>(B) * V);
  o.box_position = off;   off = up256(off + static_cast<size_t>(B) * V * 16);
  o.qa_label = off;       off = up256(off + static_cast<size_t>(B) * 8);
  o.bytes = off;
  return o;
}

struct UnpackArgs {
  const char* packed;
  Layout lay;
  int B, L, V, task;
  int64_t *word_id, *word_labels, *matched_labels, *cluster_ids, *obj_labels, *qa_labels;
  uint8_t *attention_mask, *vis_mask;
  float *additive_mask, *visual_pos;
};

// One grid-stride pass over max(B·L, B·V·4) work items; every output element is written by exactly one thread.
__global__ void unpack_kernel(const UnpackArgs a) {
  const int BL = a.B * a.L, BV = a.B * a.V;
  const int64_t* w = reinterpret_cast<const int64_t*>(a.packed + a.lay.word_id);
  const int64_t* wl = reinterpret_cast<const int64_t*>(a.packed + a.lay.word_label);
  const int64_t* ml = reinterpret_cast<const int64_t*>(a.packed + a.lay.matched_label);
  const int64_t* c = reinterpret_cast<const int64_t*>(a.packed + a.lay.cluster_id);
  const uint8_t* vm = reinterpret_cast<const uint8_t*>(a.packed + a.lay.vis_mask);
  const float4* bp = reinterpret_cast<const float4*>(a.packed + a.lay.box_position);
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BL; i += stride) {
    const int64_t id = w[i];
    a.word_id[i] = id;
    const bool keep = id > 0;                                      // attention_mask = word_id > 0   (:205)
    a.attention_mask[i] = keep ? 1 : 0;
    a.additive_mask[i] = keep ? 0.0f : -3.4028234663852886e38f;    // (1 − mask)·finfo(float32).min  (HF:766-774)
    if (a.word_labels) a.word_labels[i] = wl[i];                   // task word_mask                (:158-159)
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BV; i += stride) {
    const int64_t cid = c[i];
    const uint8_t m = vm[i] ? 1 : 0;                               // .bool()                       (:155)
    a.cluster_ids[i] = cid;
    a.vis_mask[i] = m;
    if (a.obj_labels) a.obj_labels[i] = m ? cid : kIgnore;         // obj_labels[~vis_mask] = -100  (:163-166)
    reinterpret_cast<float4*>(a.visual_pos)[i] = bp[i];
  }
  if (a.matched_labels)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += stride) a.matched_labels[i] = ml[i];
  if (a.qa_labels) {
    // --taskQA: the answer of a caption that was swapped for another image's is ignored        (:184-189)
    const int64_t* q = reinterpret_cast<const int64_t*>(a.packed + a.lay.qa_label);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += stride)
      a.qa_labels[i] = (a.task == XLX_TASK_MATCHED && ml[i] == 0) ? kIgnore : q[i];
  }
}

}  // namespace

extern "C" {

int32_t xlx_pretrain_inputs_layout(int32_t B, int32_t L, int32_t V, int64_t* offsets7, int64_t* total_bytes) {
  if (B < 1 || L < 1 || V < 1) return -21;
  if (!offsets7 || !total_bytes) return -24;
  const Layout o = layout(B, L, V);
  offsets7[0] = static_cast<int64_t>(o.word_id);
  offsets7[1] = static_cast<int64_t>(o.word_label);
  offsets7[2] = static_cast<int64_t>(o.matched_label);
  offsets7[3] = static_cast<int64_t>(o.cluster_id);
  offsets7[4] = static_cast<int64_t>(o.vis_mask);
  offsets7[5] = static_cast<int64_t>(o.box_position);
  offsets7[6] = static_cast<int64_t>(o.qa_label);
  *total_bytes = static_cast<int64_t>(o.bytes);
  return 0;
}

int32_t xlx_pretrain_inputs_unpack(const void* packed, int32_t B, int32_t L, int32_t V, int32_t task, int64_t* word_id,
                                   uint8_t* attention_mask, float* additive_mask, int64_t* cluster_ids,
                                   uint8_t* vis_mask, float* visual_pos, int64_t* obj_labels, int64_t* word_labels,
                                   int64_t* matched_labels, int64_t* qa_labels, void* stream) {
  if (B < 1 || L < 1 || V < 1) return -21;
  if (task < XLX_TASK_VIS_MASK || task > XLX_TASK_MATCHED) return -1;
  if (!packed || !word_id || !attention_mask || !additive_mask || !cluster_ids || !vis_mask || !visual_pos) return -24;
  if ((task == XLX_TASK_VIS_MASK && !obj_labels) || (task == XLX_TASK_WORD_MASK && !word_labels) ||
      (task == XLX_TASK_MATCHED && !matched_labels))
    return -24;
  if ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(visual_pos)) & 15) return -2;
  XLX_TRY(ensure_device(word_id));
  UnpackArgs a;
  a.packed = static_cast<const char*>(packed);
  a.lay = layout(B, L, V);
  a.B = B; a.L = L; a.V = V; a.task = task;
  a.word_id = word_id; a.attention_mask = attention_mask; a.additive_mask = additive_mask;
  a.cluster_ids = cluster_ids; a.vis_mask = vis_mask; a.visual_pos = visual_pos;
  a.obj_labels = task == XLX_TASK_VIS_MASK ? obj_labels : nullptr;
  a.word_labels = task == XLX_TASK_WORD_MASK ? word_labels : nullptr;
  a.matched_labels = task == XLX_TASK_MATCHED ? matched_labels : nullptr;
  a.qa_labels = qa_labels;
  const int work = B * V > B * L ? B * V : B * L;
  int blocks = (work + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  unpack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_aux_launch();
  XLX_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
