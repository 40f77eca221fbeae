/*
 * xlxmert_b200 — C ABI of the B200-native X-LXMERT hot path (libxlxmert_b200.so).
 *
 * The reference (allenai/x-lxmert) has no native layer: its hot path is Python calling torch.nn modules
 * (SURVEY.md §2.2).  The entry points below are what a Python maintainer binds with ctypes to replace the
 * *inside* of those modules; each cites the reference interface it replaces.  "HF" = the third-party
 * transformers/models/lxmert/modeling_lxmert.py (5.5.0 line numbers) that the reference imports at
 * x-lxmert/src/lxrt/modeling.py:5.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer on the calling thread's current device unless stated otherwise;
 *    float = IEEE fp32, row-major, innermost dimension contiguous.  `params` arrays are HOST arrays of
 *    device pointers.
 *  - The library never allocates, frees or retains caller memory past return; workspaces are caller-owned
 *    (size them with the *_bytes functions).  A workspace used by a forward with training=1 holds the
 *    saved activations and must be passed unchanged to the matching backward.
 *  - All work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no internal syncs.
 *  - Return value: 0 = OK; < 0 = invalid argument / unsupported shape (see xlx_strerror); > 0 = CUDA error
 *    code.  Nothing throws across the boundary and nothing is printed.
 *  - `passes`: 3 = split-bf16 "bf16x3" tensor-core GEMMs (fp32-class accuracy, the parity mode);
 *              1 = single bf16 pass (what the reference's mixed-precision flag would give, pretrain.bash:28).
 */
#ifndef XLXMERT_B200_H_
#define XLXMERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Static dimensions (HF LxmertConfig: configuration_lxmert.py:85-87 and the BERT-base sizes). */
typedef struct xlx_dims {
  int32_t hidden;        /* 768; multiple of 128, = heads * 64 */
  int32_t heads;         /* 12; head_dim must be 64 */
  int32_t intermediate;  /* 3072 */
  int32_t feat_dim;      /* 2048 */
  int32_t pos_dim;       /* 4 */
  int32_t l_layers;      /* 9  language layers    (encoder.layer)    */
  int32_t r_layers;      /* 5  vision layers      (encoder.r_layers) */
  int32_t x_layers;      /* 5  cross-modal layers (encoder.x_layers) */
  float ln_eps;          /* 1e-12 */
} xlx_dims;

const char* xlx_version(void);
const char* xlx_strerror(int32_t code);
/* number of kernels launched by this library in this process so far (bench bookkeeping) */
int64_t xlx_launch_count(void);
int64_t xlx_gemm_launch_count(void);
/* Measurement aid for bench.py's roofline: between begin and end every tcgen05 GEMM launch is bracketed by
 * CUDA events on its stream; end synchronises and returns the summed device time (ms), the summed
 * algorithmic FLOPs (2*M*N*K per launch) and the launch count. */
void xlx_profile_gemm_begin(void);
int32_t xlx_profile_gemm_end(double* total_ms, double* total_flops, int64_t* launches);

/* ---- LxmertEncoder (HF:487-565; reached from lxrt/modeling.py:195-206 via self.bert) -------------------
 * Parameter slots, in order (`params[i]` = device pointer of slot i; shapes as in the state dict):
 *   visn_fc.visn_fc.{weight[H,F],bias}, visn_fc.visn_layer_norm.{weight,bias},
 *   visn_fc.box_fc.{weight[H,4],bias}, visn_fc.box_layer_norm.{weight,bias}                     (8)
 *   then for each layer.i (i < l_layers) and each r_layers.i:   ATT(attention.self, attention.output), FFN
 *   then for each x_layers.i: ATT(visual_attention.att, visual_attention.output),
 *        ATT(lang_self_att.self, .output), ATT(visn_self_att.self, .output),
 *        FFN(lang_inter, lang_output), FFN(visn_inter, visn_output)
 *   ATT = query.weight, key.weight, value.weight, query.bias, key.bias, value.bias,
 *         output.dense.weight, output.dense.bias, output.LayerNorm.weight, output.LayerNorm.bias   (10)
 *   FFN = intermediate.dense.weight[I,H], .bias, output.dense.weight[H,I], .bias, LayerNorm.weight, .bias (6)
 * The gradient arena written by xlx_encoder_bwd is one flat fp32 buffer with the slots packed in the same
 * order (xlx_encoder_grad_offset gives each slot's element offset).
 */
int64_t xlx_encoder_num_params(const xlx_dims* d);
int64_t xlx_encoder_param_elems(const xlx_dims* d, int64_t slot);
int64_t xlx_encoder_grad_offset(const xlx_dims* d, int64_t slot);
int64_t xlx_encoder_grad_elems(const xlx_dims* d);

/* Bytes of the prepared-weights arena (bf16 hi/lo splits of every Linear weight, fused QKV). */
size_t xlx_encoder_prep_bytes(const xlx_dims* d);
/* Convert the fp32 parameters into the arena.  Call whenever the weights changed (every optimiser step). */
int32_t xlx_encoder_prepare(const xlx_dims* d, const float* const* params, void* prep, void* stream);

size_t xlx_encoder_workspace_bytes(const xlx_dims* d, int32_t B, int32_t L, int32_t V, int32_t training);

/* LxmertEncoder.forward(lang_feats, lang_attention_mask, visual_feats, visual_pos, visual_attention_mask)
 * (HF:506-565).  lang_in [B,L,H] = embedding output; lang_mask / vis_mask: additive masks [B,L] / [B,V]
 * (0 or finfo.min, HF:766-782) or NULL; visual_feats [B,V,F]; visual_pos [B,V,4].
 * Outputs: lang_out [B,L,H], vis_out [B,V,H] (last hidden states); optional per-layer hidden states
 * lang_hidden [l_layers + x_layers, B, L, H], vis_hidden [r_layers + x_layers, B, V, H] (NULL to skip). */
int32_t xlx_encoder_fwd(const xlx_dims* d, const float* const* params, const void* prep, int32_t B, int32_t L,
                        int32_t V, const float* lang_in, const float* lang_mask, const float* visual_feats,
                        const float* visual_pos, const float* vis_mask, float* lang_out, float* vis_out,
                        float* lang_hidden, float* vis_hidden, void* workspace, size_t workspace_bytes,
                        int32_t training, int32_t passes, void* stream);

/* Backward of the above (the reference gets it from torch autograd: lxmert_pretrain.py:338).
 * d_lang_out / d_vis_out: gradients wrt the two outputs (NULL = that output is not part of the loss).  Writes
 * d_lang_in [B,L,H], d_visual_feats [B,V,F] (NULL to skip) and every parameter gradient into `grads` (overwritten,
 * not accumulated).  With x_layers > 0, a NULL d_lang_out (d_vis_out) leaves the last cross-modality layer's
 * lang_self_att / lang_inter / lang_output (visn_self_att / visn_inter / visn_output) parameters without gradient, as
 * in the reference's autograd graph: those blocks are skipped and their 16 arena slots are NOT written — the caller
 * must report "no gradient" for them (the reference loop's `param.grad = None`, lxmert_pretrain.py:363-364).
 * `stages` is a mask of XLX_BWD_* (XLX_BWD_ALL for the whole backward).  The stages must be issued in the order
 * CROSS, VISION, LANGUAGE, VISN_FC, in one call or several (the state between stages lives in the workspace): a
 * data-parallel caller enqueues the all-reduce of the gradient-arena range a stage completed
 * (xlx_encoder_grad_stage_range) while the next stage computes — the overlap DistributedDataParallel gets from its
 * bucket hooks (lxmert_pretrain.py:102-106), with four static buckets instead of an autograd-graph walk. */
enum {
  XLX_BWD_CROSS = 1,    /* x_layers (last in the forward, first in the backward) */
  XLX_BWD_VISION = 2,   /* r_layers */
  XLX_BWD_LANGUAGE = 4, /* layer.* and d_lang_in */
  XLX_BWD_VISN_FC = 8,  /* visn_fc and d_visual_feats */
  XLX_BWD_ALL = 15
};
int32_t xlx_encoder_bwd(const xlx_dims* d, const float* const* params, const void* prep, int32_t B, int32_t L,
                        int32_t V, const float* visual_pos, const float* d_lang_out, const float* d_vis_out,
                        float* d_lang_in, float* d_visual_feats, float* grads, void* workspace,
                        size_t workspace_bytes, int32_t passes, int32_t stages, void* stream);
int32_t xlx_encoder_grad_stage_range(const xlx_dims* d, int32_t stage, int64_t* offset, int64_t* elems);

/* ---- visual input of XLxmertForPretraining.forward (x-lxmert/src/lxrt/modeling.py:185-193) ----------------------
 * out[r,:] = vis_mask[r] ? mask_feat[:] : table[cluster_ids[r],:]   (vis_emb gather + torch.where), r < rows.
 * vis_mask: one byte per row (torch.bool) or NULL.  Backward: d_mask_feat[:] = Σ_{r: vis_mask[r]} d_feats[r,:];
 * the table is frozen (nn.Embedding.from_pretrained(freeze=True), modeling.py:146-149).
 * scratch: 128·feat_dim floats. */
int32_t xlx_visual_input_fwd(const float* table, const int64_t* cluster_ids, const uint8_t* vis_mask,
                             const float* mask_feat, int32_t rows, int32_t feat_dim, float* out, void* stream);
int32_t xlx_visual_input_bwd(const float* d_feats, const uint8_t* vis_mask, int32_t rows, int32_t feat_dim,
                             float* d_mask_feat, float* scratch, void* stream);

/* ---- input path of the pre-training step --------------------------------------------------
 * Device half of Trainer.forward (x-lxmert/src/pretrain/lxmert_pretrain.py:143-225): the host packs the arrays of one
 * collate_fn batch (lxmert_data.py:497-652) into ONE pinned buffer with the layout below, copies it with one
 * cudaMemcpyAsync, and this kernel unpacks it into the tensors XLxmertForPretraining.forward takes.
 *   layout (byte offsets, each section 256-byte aligned; xlx_pretrain_inputs_layout fills offsets6 in this order):
 *     [0] word_id       int64 [B,L]  — the ids the task reads: masked_word_id (word_mask), other_word_id (matched),
 *                                      word_id (vis_mask)                                     (:193-198)
 *     [1] word_label    int64 [B,L]  — read for task word_mask only                            (:158-159)
 *     [2] matched_label int64 [B]    — read for task matched only                              (:181-183)
 *     [3] cluster_id    int64 [B,V]                                                            (:147-149)
 *     [4] vis_mask      uint8 [B,V]                                                            (:155)
 *     [5] box_position  fp32  [B,V,4]                                                          (:153)
 *   outputs: word_id, attention_mask = word_id > 0 (one byte per token, torch.bool), additive_mask [B,L] =
 *     (1 − mask)·finfo.min (HF:766-774; what xlx_encoder_fwd takes), cluster_ids, vis_mask (bool bytes), visual_pos,
 *     and the task's labels: obj_labels = vis_mask ? cluster_id : −100 (:163-166), word_labels, matched_labels
 *     (the other two may be NULL). */
enum { XLX_TASK_VIS_MASK = 0, XLX_TASK_WORD_MASK = 1, XLX_TASK_MATCHED = 2 };   /* MASK_MODALITY order, :794-800 */
int32_t xlx_pretrain_inputs_layout(int32_t B, int32_t L, int32_t V, int64_t* offsets6, int64_t* total_bytes);
int32_t xlx_pretrain_inputs_unpack(const void* packed, int32_t B, int32_t L, int32_t V, int32_t task, int64_t* word_id,
                                   uint8_t* attention_mask, float* additive_mask, int64_t* cluster_ids,
                                   uint8_t* vis_mask, float* visual_pos, int64_t* obj_labels, int64_t* word_labels,
                                   int64_t* matched_labels, void* stream);

/* ---- LxmertEmbeddings (HF:179-214; reached through self.bert at lxrt/modeling.py:195-206) -----------------
 * params (HOST array of 5 device pointers): word_embeddings.weight [vocab,H], position_embeddings.weight
 * [max_pos,H], token_type_embeddings.weight [type_vocab,H], LayerNorm.weight, LayerNorm.bias.
 * out[b,l,:] = LN(word[ids[b,l]] + pos[l] + type[tt[b,l]]); dropout is the caller's business (p = 0 / eval).
 * input_ids / token_type_ids: int64 [B,L] (token_type_ids may be NULL = all zero, as in every reference caller).
 * `save` (xlx_embeddings_save_bytes) receives what the backward needs; NULL for inference. */
size_t xlx_embeddings_save_bytes(const xlx_dims* d, int32_t B, int32_t L);
size_t xlx_embeddings_scratch_bytes(const xlx_dims* d, int32_t B, int32_t L);
int32_t xlx_embeddings_fwd(const xlx_dims* d, int32_t B, int32_t L, const int64_t* input_ids,
                           const int64_t* token_type_ids, const float* const* params, float* out, void* save,
                           void* stream);
/* grads: HOST array of 5 device pointers shaped like params, all overwritten.  Row 0 of each table gets no
 * gradient (nn.Embedding(padding_idx=0), HF:184-186). */
int32_t xlx_embeddings_bwd(const xlx_dims* d, int32_t B, int32_t L, int32_t vocab, int32_t max_pos,
                           int32_t type_vocab, const int64_t* input_ids, const int64_t* token_type_ids,
                           const float* const* params, const void* save, const float* d_out, float* const* grads,
                           void* scratch, size_t scratch_bytes, void* stream);

/* ---- LxmertPooler (HF:568-580): pooled = tanh(W · lang_out[:,0] + b) ---------------------------------------
 * The workspace written by the forward must reach the backward unchanged.  d_lang_out [B,L,H] is fully
 * written (zero except token 0 of every sample). */
size_t xlx_pooler_workspace_bytes(const xlx_dims* d, int32_t B);
int32_t xlx_pooler_fwd(const xlx_dims* d, int32_t B, int32_t L, const float* lang_out, const float* W,
                       const float* bias, float* pooled, void* workspace, size_t workspace_bytes, int32_t passes,
                       void* stream);
int32_t xlx_pooler_bwd(const xlx_dims* d, int32_t B, int32_t L, const float* pooled, const float* d_pooled,
                       float* d_lang_out, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                       int32_t passes, void* stream);

/* ---- prediction heads ---------------------------------------------------------------------------------------
 * Cluster head = lxrt LxmertVisualObjHead (x-lxmert/src/lxrt/modeling.py:8-53, cluster mode):
 *     t = LN(gelu(W1·h + b1));  feat = Wf·t + bf  [M, feat_dim];  obj = feat·Centroidsᵀ + bc  [M, classes]
 *   params (8): transform.dense.{weight,bias}, transform.LayerNorm.{weight,bias}, linear_feat.{weight,bias},
 *               out_cluster.{weight [classes,feat_dim] (the frozen centroid table, modeling.py:146-151), bias}
 * LM head = HF LxmertLMPredictionHead (HF:597-607) as used by LxmertPreTrainingHeads (HF:656-665):
 *     t = LN(gelu(W1·h + b1));  scores = t·Eᵀ + bias  [M, vocab]
 *   params (6): predictions.transform.dense.{weight,bias}, predictions.transform.LayerNorm.{weight,bias},
 *               predictions.decoder.weight [vocab,H] (tied to the word embeddings, modeling.py:86), predictions.bias
 * `prep` holds split-bf16 weight copies (refresh with *_prepare when weights change).  hidden: [M,H] fp32.
 * Optional outputs (NULL to skip): feat [M,feat_dim], logits/scores [M,classes], loss (device scalar =
 * CrossEntropyLoss() over rows with label != -100, modeling.py:99,253-256; needs labels), and for the sampler
 * pred_prob [M] / pred_id [M] = softmax(logits).max(-1) with torch's first-index tie rule
 * (tasks/imggen_model.py:232-235).  The workspace written by a forward with labels must reach the backward
 * unchanged.  Backward: d_loss = device scalar (upstream gradient of the loss); writes d_hidden [M,H] and
 * grads (HOST array of device pointers shaped like params; out_cluster.weight is frozen and not written). */
size_t xlx_objhead_prep_bytes(const xlx_dims* d, int32_t classes);
int32_t xlx_objhead_prepare(const xlx_dims* d, int32_t classes, const float* const* params, void* prep, void* stream);
size_t xlx_objhead_workspace_bytes(const xlx_dims* d, int32_t classes, int32_t M);
int32_t xlx_objhead_fwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                        const float* hidden, const int64_t* labels, float* feat, float* logits, float* loss,
                        float* pred_prob, int64_t* pred_id, void* workspace, size_t workspace_bytes, int32_t passes,
                        void* stream);
int32_t xlx_objhead_bwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                        const int64_t* labels, const float* d_loss, float* d_hidden, float* const* grads,
                        void* workspace, size_t workspace_bytes, int32_t passes, void* stream);
size_t xlx_lmhead_prep_bytes(const xlx_dims* d, int32_t classes);
int32_t xlx_lmhead_prepare(const xlx_dims* d, int32_t classes, const float* const* params, void* prep, void* stream);
size_t xlx_lmhead_workspace_bytes(const xlx_dims* d, int32_t classes, int32_t M);
int32_t xlx_lmhead_fwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                       const float* hidden, const int64_t* labels, float* scores, float* loss, void* workspace,
                       size_t workspace_bytes, int32_t passes, void* stream);
int32_t xlx_lmhead_bwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                       const int64_t* labels, const float* d_loss, float* d_hidden, float* const* grads,
                       void* workspace, size_t workspace_bytes, int32_t passes, void* stream);

/* Row compaction around the two masked-prediction losses.  Replaces nothing in the reference: CrossEntropyLoss
 * (lxrt/modeling.py:99,102 and :253-256) ignores rows labelled -100, yet the reference still runs the heads on them.
 * rows[0..*count) = ascending m with labels[m] != ignore_index (rows holds M entries; count is a DEVICE int32 the
 * caller reads back to size the head call).  gather: dst[i,:] = src[rows[i],:], labels_dst[i] = labels[rows[i]]
 * (either pair src/dst or labels/labels_dst may be null).  scatter: dst [M, cols] = 0 then dst[rows[i],:] = src[i,:] — the gather's adjoint. */
int32_t xlx_labelled_rows(const int64_t* labels, int32_t M, int64_t ignore_index, int64_t* rows, int32_t* count,
                          void* stream);
int32_t xlx_gather_rows(const float* src, const int64_t* labels, const int64_t* rows, int32_t n, int32_t cols,
                        float* dst, int64_t* labels_dst, void* stream);
int32_t xlx_scatter_rows(const float* src, const int64_t* rows, int32_t n, int32_t M, int32_t cols, float* dst,
                         void* stream);

/* Matched head: cls.seq_relationship = Linear(H, 2) on the pooled output (HF:661,664) + CrossEntropyLoss
 * (modeling.py:227-235).  scores [B,2]; scratch = xlx_matchhead_scratch_floats(B) floats, shared by fwd and bwd. */
int64_t xlx_matchhead_scratch_floats(int32_t B);
int32_t xlx_matchhead_fwd(const xlx_dims* d, int32_t B, const float* pooled, const float* W, const float* bias,
                          const int64_t* labels, float* scores, float* loss, float* scratch, void* stream);
int32_t xlx_matchhead_bwd(const xlx_dims* d, int32_t B, const float* pooled, const float* W, const int64_t* labels,
                          const float* scores, const float* d_loss, float* d_pooled, float* dW, float* dbias,
                          float* scratch, void* stream);

/* ---- sampling-loop transitions (tasks/imggen_model.py:49-257) ----------------------
 * The index / select work between two encoder passes, on the device (the reference: torch.topk + scatter_ +
 * torch.where + nn.Embedding).  code [B,V,F] fp32 (in place), vis_mask / visited: one byte per cell, pred_prob [B,V],
 * pred_id [B,V] from xlx_objhead_fwd, table = centroid table [classes,F], mask_feat [F].  V ≤ 64, F % 4 == 0.
 * Ties between EQUAL probabilities go to the lower cell index (torch.topk leaves that order unspecified).
 *   nar_update  (mask-predict, :199-243): pred_prob == NULL → initial state (all cells masked, code = mask_feat);
 *               else code ← vis_mask ? table[pred_id] : code  (:238-243), vis_mask_next ← the n_mask_next cells of lowest
 *               pred_prob (:209-212), code ← vis_mask_next ? mask_feat : code (:215-218 of the next iteration).
 *               vis_mask_next may alias vis_mask.
 *   ar_update   (one cell per step, :135-153): position ≥ 0 → that cell; < 0 → the unvisited cell of highest pred_prob;
 *               code[b,cell] ← table[pred_id[b,cell]], vis_mask ← 0 there, visited ← 1 (confidence order only).
 *   remask_cell (:111-113): code[:,cell] ← mask_feat, vis_mask[:,cell] ← 1. */
int32_t xlx_sampler_nar_update(float* code, const uint8_t* vis_mask, const float* pred_prob, const int64_t* pred_id,
                               const float* table, const float* mask_feat, int32_t B, int32_t V, int32_t F,
                               int32_t n_mask_next, uint8_t* vis_mask_next, void* stream);
int32_t xlx_sampler_ar_update(float* code, uint8_t* vis_mask, uint8_t* visited, const float* pred_prob,
                              const int64_t* pred_id, const float* table, int32_t B, int32_t V, int32_t F,
                              int32_t position, void* stream);
int32_t xlx_sampler_remask_cell(float* code, uint8_t* vis_mask, const float* mask_feat, int32_t B, int32_t V, int32_t F,
                                int32_t cell, void* stream);

/* ---- Generator.forward (image_generator/src/layers.py:223-253) ----------------------------------------------
 * Canonical architecture only (scripts/train_generator.bash, tasks/sample_images.py:53-67): base_dim 32, emb_dim
 * 2048, codebook_dim 256, norm 'spade_in', SN, 8×8 grid → 256×256 RGB, five up-sampling residual blocks.
 * params: HOST array of xlx_generator_num_params() = 150 device pointers in this order:
 *   bottleneck_emb.0.{weight,bias}; learned_init_conv.0.{weight_orig,bias,weight_u,weight_v}; style_init_conv.0.{…};
 *   per block i: cbn1.{shared.0,gamma,beta}.{weight,bias}, cbn2.{…}, conv1.{weight_orig,bias,weight_u,weight_v},
 *   conv2.{…}, res_branch.1.{…}, noise1.weight, noise2.weight;  then to_RGB_blocks.i.conv.{weight,bias}.
 * Spectral norm has eval semantics: weight = weight_orig / (uᵀ·W·v) with the stored u, v (folded by *_prepare).
 * emb: [B, 64, 2048] fp32, grid-cell major (the memory layout behind the sampler's
 * code.permute(0,2,1).view(B,2048,8,8), tasks/imggen_model.py:254, and layers.py:231-233's [B,8,8,2048]).
 * noise: NULL (forward(train=False)) or HOST array of 10 device pointers, [B,R,R] standard-normal maps for
 * noise1 / noise2 of each block (R = 8,16, 16,32, …, 128,256; layers.py:56-62).
 * img: [B,3,256,256] fp32 (NCHW like the reference); pre_tanh (optional) same shape; block_out (optional): HOST
 * array of 5 device pointers receiving each block's output as NHWC [B,2R,2R,32]. */
int64_t xlx_generator_num_params(void);
int64_t xlx_generator_launch_count(void);
size_t xlx_generator_prep_bytes(void);
int32_t xlx_generator_prepare(const float* const* params, void* prep, void* stream);
size_t xlx_generator_workspace_bytes(int32_t B);
int32_t xlx_generator_fwd(const float* const* params, const void* prep, int32_t B, const float* emb,
                          const float* const* noise, float* img, float* pre_tanh, float* const* block_out,
                          void* workspace, size_t workspace_bytes, int32_t passes, void* stream);

/* ---- nearest-centroid assignment -----------------------------------------------------------
 * x-lxmert/feature_extraction/run_kmeans.py:124-143: index = faiss.IndexFlatL2(d); index.add(centroids);
 * D, I = index.search(x, 1)  — the cluster id of every grid cell.  faiss 1.6 (not vendored) computes
 * ‖x‖² + ‖c‖² − 2·x·c with an SGEMM and keeps the minimum; the same arithmetic here, as one tcgen05 GEMM against the
 * prepared table split-bf16(2·C) with −‖c‖² as bias, followed by a row arg-max (ties → lowest index).
 *   prepare: centroids [n_centroids, dim] fp32 → prep (xlx_kmeans_prep_bytes bytes); redo when the table changes.
 *   assign:  x [N, dim] fp32 → ids [N] int64 and (optional) dist [N] fp32 = squared L2 distance to that centroid,
 *            clamped at 0 like faiss.  workspace: xlx_kmeans_workspace_bytes(dim, n_centroids, N). dim % 8 == 0. */
size_t xlx_kmeans_prep_bytes(int32_t dim, int32_t n_centroids);
int32_t xlx_kmeans_prepare(int32_t dim, int32_t n_centroids, const float* centroids, void* prep, void* stream);
size_t xlx_kmeans_workspace_bytes(int32_t dim, int32_t n_centroids, int32_t N);
int32_t xlx_kmeans_assign(int32_t dim, int32_t n_centroids, const void* prep, int32_t N, const float* x, int64_t* ids,
                          float* dist, void* workspace, size_t workspace_bytes, int32_t passes, void* stream);

/* ---- optimiser step of the pre-training loop ---------------------------------------------
 * x-lxmert/src/pretrain/lxmert_pretrain.py:343-364: torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) then
 * transformers.optimization.AdamW.step() (4.1.1: betas (0.9, 0.999), eps 1e-6, correct_bias, weight decay applied
 * after the Adam update as p -= lr·wd·p; :110-141 builds the two decay groups).  All arrays are HOST arrays of n
 * entries (device pointers / element counts / per-tensor weight decay); tensors must be 16-byte aligned.
 *   xlx_grad_sqnorm: *out (device scalar) = Σ over all tensors of ‖g‖² (deterministic two-stage reduction);
 *                    scratch: xlx_optim_scratch_floats(elems, n) floats.
 *   xlx_adamw_step:  in-place update of params / exp_avg / exp_avg_sq; `step` is the 1-based update count.
 *                    sqnorm (device scalar from xlx_grad_sqnorm, or NULL) with max_grad_norm > 0 applies the clip
 *                    coefficient min(1, max_norm / (sqrt(sqnorm) + 1e-6)) to the gradients on the fly. */
int64_t xlx_optim_scratch_floats(const int64_t* elems, int32_t n);
int64_t xlx_optim_launch_count(void);
int32_t xlx_grad_sqnorm(const float* const* grads, const int64_t* elems, int32_t n, float* scratch, float* out,
                        void* stream);
int32_t xlx_adamw_step(float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* elems, const float* weight_decay, int32_t n, double lr,
                       double beta1, double beta2, double eps, int32_t step, int32_t correct_bias, const float* sqnorm,
                       double max_grad_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XLXMERT_B200_H_ */
