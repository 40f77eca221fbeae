// C ABI of the pieces around the encoder: LxmertEmbeddings (HF modeling_lxmert.py:179-214), LxmertPooler
// (HF:568-580), and the pre-training heads — the reference's cluster head lxrt/modeling.py:8-53 with its
// cross-entropy (:244-258) / sampler arg-max (tasks/imggen_model.py:232-235), and HF's LM head (HF:597-607,656-665).
#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

using namespace xlx;

namespace {

bool hidden_ok(const xlx_dims* d) { return d && d->hidden > 0 && d->hidden % 128 == 0 && d->hidden <= 1024; }

struct EmbSave { float *y, *mean, *rstd; size_t bytes; };
EmbSave emb_layout(const xlx_dims* d, int M, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  EmbSave s;
  s.y = b.f32(static_cast<size_t>(M) * d->hidden);
  s.mean = b.f32(M);
  s.rstd = b.f32(M);
  s.bytes = b.total();
  return s;
}

struct PoolWs { Split x0, w, dpre; size_t bytes; };
PoolWs pool_layout(const xlx_dims* d, int B, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  const size_t H = d->hidden;
  PoolWs w;
  w.x0 = b.split(B * H);
  w.w = b.split(H * H);
  w.dpre = b.split(B * H);
  w.bytes = b.total();
  return w;
}


// ---- prediction heads ---------------------------------------------------------------------------------
// All three heads are  t = LN(gelu(W1·h + b1))  followed by a classifier:
//   cluster head:  W1: H→H [LxmertPredictionHeadTransform, HF:583-594]; feat = Wf·t + bf (H → F);
//                  logits = feat·Centroidsᵀ + bc (F → C)                                     lxrt/modeling.py:38-53
//   LM head:       W1: H→H; logits = t·Eᵀ + bias (H → vocab, E tied to the word embeddings)   HF:597-607
//   answer head:   W1: H→2H, LN over 2H; logits = W2·t + b2 (2H → answers)                    HF:610-623
// T = width of the transform (H or 2H); `F == 0` selects the forms without the feature layer.  Cp = C rounded up to a
// multiple of 8 (GEMM epilogue vector width).
struct HeadDims { int H, T, F, C, Cp; float eps; };
enum HeadKind { HEAD_LM = 0, HEAD_CLUSTER = 1, HEAD_ANSWER = 2 };
inline int pad8(int c) { return (c + 7) & ~7; }

struct HeadPrep { Split w1, wf, wc; float* bc; size_t bytes; };
HeadPrep head_prep_layout(const HeadDims& h, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  HeadPrep p;
  p.w1 = b.split(static_cast<size_t>(h.T) * h.H);
  if (h.F) p.wf = b.split(static_cast<size_t>(h.F) * h.T);
  p.wc = b.split(static_cast<size_t>(h.C) * (h.F ? h.F : h.T));
  p.bc = b.f32(h.Cp);                                   // class bias, zero padded to Cp
  p.bytes = b.total();
  return p;
}

struct HeadWs {
  Split x, t, feat, dlogits, dfeat, du;
  float *u, *g, *mean, *rstd, *feat32, *featrow, *logits, *lse, *rowloss, *stats, *dt, *dg, *part, *dbc, *splitk, *rowstat;
  size_t bytes;
};
HeadWs head_ws_layout(const HeadDims& h, int M, void* base) {
  Bump b; b.base = static_cast<char*>(base);
  HeadWs w;
  const size_t m = M, H = h.H, T = h.T, F = h.F, Cp = h.Cp;
  w.x = b.split(m * H);
  w.u = b.f32(m * T);
  w.g = b.f32(m * T);
  w.t = b.split(m * T);
  w.mean = b.f32(m);
  w.rstd = b.f32(m);
  w.feat32 = w.featrow = nullptr;
  if (F) {
    w.feat = b.split(m * F);
    w.feat32 = b.f32(m * F);      // fp32 features for the regression loss; its gradient overwrites them in the backward
    w.featrow = b.f32(m);
  }
  w.logits = b.f32(m * Cp);
  w.lse = b.f32(m);
  w.rowloss = b.f32(m);
  w.stats = b.f32(8);             // [0] CE loss [1] valid rows [4] feature-regression loss
  // backward scratch
  w.dlogits = b.split(m * Cp);
  if (F) w.dfeat = b.split(m * F);
  w.dt = b.f32(m * T);
  w.dg = b.f32(m * T);
  w.du = b.split(m * T);
  size_t pe = 3 * static_cast<size_t>(reduce_max_blocks()) * T;
  const size_t widest = Cp > F ? Cp : F;
  if (128 * widest > pe) pe = 128 * widest;
  w.part = b.f32(pe);
  w.dbc = b.f32(Cp);
  w.splitk = b.f32(gemm_splitk_ws_floats());
  w.rowstat = b.f32(m * gemm_rowstat_slots(h.Cp) * 3);
  w.bytes = b.total();
  return w;
}

// params: [0] first Linear weight [T,H] [1] its bias [2] LayerNorm.weight [T] [3] LayerNorm.bias, then
//   cluster head: [4] linear_feat.weight [5] .bias [6] out_cluster.weight [7] out_cluster.bias
//   LM head:      [4] decoder.weight [5] predictions.bias
//   answer head:  [4] logit_fc.3.weight [5] logit_fc.3.bias
int head_prepare(const HeadDims& h, const float* const* params, void* prep, cudaStream_t st) {
  HeadPrep p = head_prep_layout(h, prep);
  const size_t H = h.H, T = h.T;
  XLX_TRY(split_f32(params[0], p.w1, T * H, st));
  const float* wc = params[h.F ? 6 : 4];
  const float* bc = params[h.F ? 7 : 5];
  if (h.F) XLX_TRY(split_f32(params[4], p.wf, static_cast<size_t>(h.F) * T, st));
  XLX_TRY(split_f32(wc, p.wc, static_cast<size_t>(h.C) * (h.F ? h.F : T), st));
  XLX_CUDA(cudaMemsetAsync(p.bc, 0, static_cast<size_t>(h.Cp) * 4, st));
  XLX_CUDA(cudaMemcpyAsync(p.bc, bc, static_cast<size_t>(h.C) * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Feature-regression loss inputs of the cluster head (lxrt/modeling.py:270-284); all null = loss off
struct FeatLoss {
  const float* target = nullptr;   // [M, F] feat_labels rows
  const float* weight = nullptr;   // [M] from xlx_feat_row_weight
  float* loss = nullptr;           // forward: device scalar out
  const float* d_loss = nullptr;   // backward: device scalar in
};

int head_fwd(const HeadDims& h, const float* const* params, const void* prep_base, int M, const float* hidden,
             const int64_t* labels, FeatLoss fl, float* feat_out, float* logits_out, float* loss, float* pred_prob,
             int64_t* pred_id, void* ws_base, size_t ws_bytes, int passes, cudaStream_t st) {
  HeadPrep p = head_prep_layout(h, const_cast<void*>(prep_base));
  HeadWs w = head_ws_layout(h, M, ws_base);
  if (w.bytes > ws_bytes) return -23;
  const int H = h.H, T = h.T, F = h.F, C = h.C, Cp = h.Cp;
  if (fl.target && (!F || !fl.weight)) return -24;
  XLX_TRY(split_f32(hidden, w.x, static_cast<size_t>(M) * H, st));
  {
    GemmEpilogue e;   // u = W1·h + b1 (saved), g = gelu(u)
    e.bias = params[1]; e.flags = EPI_GELU; e.out_u = w.u; e.ld_u = T; e.out_f32 = w.g; e.ld_out = T;
    XLX_TRY(gemm_linear(passes, st, w.x, M, H, p.w1, T, e));
  }
  XLX_TRY(layernorm_fwd(w.g, params[2], params[3], h.eps, M, T, 1.0f, nullptr, w.t, nullptr, w.mean, w.rstd, st));
  Split dec_in = w.t;
  int Kc = T;
  if (F) {
    GemmEpilogue e;
    e.bias = params[5]; e.out_f32 = fl.target ? w.feat32 : feat_out; e.ld_out = F;
    e.out_hi = w.feat.hi; e.out_lo = w.feat.lo; e.ld_split = F;
    XLX_TRY(gemm_linear(passes, st, w.t, M, T, p.wf, F, e));
    if (fl.target) {
      const size_t bytes = static_cast<size_t>(M) * F * 4;
      if (feat_out) XLX_CUDA(cudaMemcpyAsync(feat_out, w.feat32, bytes, cudaMemcpyDeviceToDevice, st));
      XLX_TRY(smooth_l1_fwd(w.feat32, fl.target, fl.weight, M, F, w.featrow, w.stats + 4, st));
      if (fl.loss) XLX_CUDA(cudaMemcpyAsync(fl.loss, w.stats + 4, 4, cudaMemcpyDeviceToDevice, st));
    }
    dec_in = w.feat; Kc = F;
  }
  if (!labels && !logits_out && !pred_prob) return 0;     // features only (--visualLosses feat)
  if (pred_prob && pred_id && !logits_out && !labels) {
    // sampler step (tasks/imggen_model.py:228-235): softmax(logits).max(-1) straight from the accumulators — the
    // [M, classes] logits are never written; per-tile partial statistics (12 B per row and quarter tile) are merged
    static const bool fused = [] { const char* e = getenv("XLX_FUSED_ARGMAX"); return !(e && e[0] == '0'); }();
    if (fused && !getenv("XLX_GEMM_BN")) {
      GemmEpilogue e;
      e.bias = p.bc; e.rowstat = w.rowstat; e.rowstat_cols = C;
      XLX_TRY(gemm_linear(passes, st, dec_in, M, Kc, p.wc, Cp, e, C));
      return rowstat_merge(w.rowstat, M, gemm_rowstat_slots(Cp), pred_prob, pred_id, st);
    }
  }
  {
    GemmEpilogue e;
    e.bias = p.bc; e.out_f32 = w.logits; e.ld_out = Cp;
    XLX_TRY(gemm_linear(passes, st, dec_in, M, Kc, p.wc, Cp, e, C));
  }
  if (logits_out)
    XLX_CUDA(cudaMemcpy2DAsync(logits_out, static_cast<size_t>(C) * 4, w.logits, static_cast<size_t>(Cp) * 4,
                               static_cast<size_t>(C) * 4, M, cudaMemcpyDeviceToDevice, st));
  if (labels) {
    XLX_TRY(ce_fwd(w.logits, Cp, M, C, labels, -100, w.lse, w.rowloss, w.stats, st));
    if (loss) XLX_CUDA(cudaMemcpyAsync(loss, w.stats, 4, cudaMemcpyDeviceToDevice, st));
  }
  if (pred_prob && pred_id) XLX_TRY(softmax_argmax(w.logits, Cp, M, C, pred_prob, pred_id, st));
  return 0;
}

// grads: device pointers shaped like params (overwritten; a null entry is skipped).  The cluster head's
// out_cluster.weight ([6]) is the frozen centroid table (lxrt/modeling.py:146-151) and gets no gradient; the LM head's
// decoder.weight ([4]) does.  The logits gradient comes from the fused cross-entropy (labels + d_loss) or from the caller
// (d_logits [M, C], a differentiable `forward`); the cluster head may instead / also receive a feature gradient: the
// regression loss (fl) or an upstream d_feat [M, F] — not both.
int head_bwd(const HeadDims& h, const float* const* params, const void* prep_base, int M, const int64_t* labels,
             const float* d_loss, const float* d_logits, FeatLoss fl, const float* d_feat, float* d_hidden,
             float* const* grads, void* ws_base, size_t ws_bytes, int passes, cudaStream_t st) {
  HeadPrep p = head_prep_layout(h, const_cast<void*>(prep_base));
  HeadWs w = head_ws_layout(h, M, ws_base);
  if (w.bytes > ws_bytes) return -23;
  const int H = h.H, T = h.T, F = h.F, C = h.C, Cp = h.Cp;
  if (labels && (d_logits || !d_loss)) return -24;
  if (fl.target && (!F || !fl.weight || !fl.d_loss || d_feat)) return -24;
  if (d_feat && !F) return -24;
  const bool have_dlogits = labels || d_logits;
  if (!have_dlogits && !fl.target && !d_feat) return -24;
  XLX_TRY(gemm_splitk_ws_reset(w.splitk, st));
  if (labels) XLX_TRY(ce_bwd(w.logits, Cp, M, C, Cp, labels, -100, w.lse, w.stats, d_loss, w.dlogits, st));
  else if (d_logits) XLX_TRY(split_pad_f32(d_logits, M, C, Cp, w.dlogits, st));
  float* g_cbias = grads[F ? 7 : 5];
  if (have_dlogits && g_cbias) {
    // class bias: column sums over the padded width into scratch, first C entries are the gradient
    XLX_TRY(colsum(nullptr, w.dlogits, M, Cp, Cp, w.part, w.dbc, st));
    XLX_CUDA(cudaMemcpyAsync(g_cbias, w.dbc, static_cast<size_t>(C) * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (F) {
    const float* addend = d_feat;
    if (fl.target) {
      XLX_TRY(smooth_l1_bwd(w.feat32, fl.target, fl.weight, fl.d_loss, M, F, w.feat32, st));
      addend = w.feat32;
    }
    if (have_dlogits) {
      GemmEpilogue e;   // dfeat = dlogits · Centroids (+ the feature gradient)
      e.out_hi = w.dfeat.hi; e.out_lo = w.dfeat.lo; e.ld_split = F; e.addend = addend; e.ld_addend = F;
      XLX_TRY(gemm_dgrad(passes, st, w.dlogits, M, Cp, p.wc, F, e, C));
    } else {
      XLX_TRY(split_f32(addend, w.dfeat, static_cast<size_t>(M) * F, st));
    }
    XLX_TRY(colsum(nullptr, w.dfeat, M, F, F, w.part, grads[5], st));
    XLX_TRY(gemm_wgrad(passes, st, w.dfeat, M, F, w.t, T, grads[4], false, 0, w.splitk));
    GemmEpilogue o;
    o.out_f32 = w.dt; o.ld_out = T;
    XLX_TRY(gemm_dgrad(passes, st, w.dfeat, M, F, p.wf, T, o));
  } else {
    XLX_TRY(gemm_wgrad(passes, st, w.dlogits, M, C, w.t, T, grads[4], false, Cp, w.splitk));   // dW [classes, T]
    GemmEpilogue o;
    o.out_f32 = w.dt; o.ld_out = T;
    XLX_TRY(gemm_dgrad(passes, st, w.dlogits, M, Cp, p.wc, T, o, C));
  }
  int nblk = 0;
  XLX_TRY(layernorm_bwd(w.dt, 1.0f, w.g, params[2], w.mean, w.rstd, M, T, w.dg, Split(), w.part, &nblk, st));
  float* o2[2] = {grads[2], grads[3]};
  XLX_TRY(colsum_finish(w.part, 2, nblk, T, o2, 0, st));
  XLX_TRY(gelu_bwd_split(w.dg, w.u, w.du, static_cast<size_t>(M) * T, st));
  XLX_TRY(colsum(nullptr, w.du, M, T, T, w.part, grads[1], st));
  XLX_TRY(gemm_wgrad(passes, st, w.du, M, T, w.x, H, grads[0], false, 0, w.splitk));
  GemmEpilogue o;
  o.out_f32 = d_hidden; o.ld_out = H;
  return gemm_dgrad(passes, st, w.du, M, T, p.w1, H, o);
}

bool head_dims(const xlx_dims* d, int classes, int kind, HeadDims* h) {
  if (!hidden_ok(d) || classes < 1) return false;
  if (kind == HEAD_CLUSTER && (d->feat_dim < 8 || d->feat_dim % 8)) return false;
  h->H = d->hidden; h->T = kind == HEAD_ANSWER ? 2 * d->hidden : d->hidden;
  h->F = kind == HEAD_CLUSTER ? d->feat_dim : 0; h->C = classes; h->Cp = pad8(classes); h->eps = d->ln_eps;
  return true;
}

}  // namespace

extern "C" {

// ---- visual input (lxrt/modeling.py:185-193) --------------------------------------------------------
int32_t xlx_visual_input_fwd(const float* table, const int64_t* cluster_ids, const uint8_t* vis_mask,
                             const float* mask_feat, int32_t rows, int32_t feat_dim, float* out, void* stream) {
  if (rows < 1 || feat_dim < 4) return -21;
  if (!table || !cluster_ids || !out || (vis_mask && !mask_feat)) return -24;
  XLX_TRY(ensure_device(out));
  return gather_rows(table, cluster_ids, vis_mask, mask_feat, rows, feat_dim, out, Split(),
                     static_cast<cudaStream_t>(stream));
}
int32_t xlx_visual_input_bwd(const float* d_feats, const uint8_t* vis_mask, int32_t rows, int32_t feat_dim,
                             float* d_mask_feat, float* scratch, void* stream) {
  if (rows < 1 || feat_dim < 4) return -21;
  if (!d_feats || !vis_mask || !d_mask_feat || !scratch) return -24;
  XLX_TRY(ensure_device(d_mask_feat));
  return colsum(d_feats, Split(), rows, feat_dim, feat_dim, scratch, d_mask_feat, static_cast<cudaStream_t>(stream),
                vis_mask);
}

// ---- embeddings ---------------------------------------------------------------------------------
size_t xlx_embeddings_save_bytes(const xlx_dims* d, int32_t B, int32_t L) {
  return (hidden_ok(d) && B > 0 && L > 0) ? emb_layout(d, B * L, nullptr).bytes : 0;
}
size_t xlx_embeddings_scratch_bytes(const xlx_dims* d, int32_t B, int32_t L) {
  if (!hidden_ok(d) || B < 1 || L < 1) return 0;
  return (static_cast<size_t>(B) * L * d->hidden + 3 * static_cast<size_t>(reduce_max_blocks()) * d->hidden) * 4 + 512;
}

int32_t xlx_embeddings_fwd(const xlx_dims* d, int32_t B, int32_t L, const int64_t* input_ids,
                           const float* inputs_embeds, const int64_t* token_type_ids, const float* const* params,
                           float* out, void* save, const xlx_dropout* dropout, void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1 || L < 1) return -21;
  if ((!input_ids == !inputs_embeds) || !params || !out) return -24;     // exactly one of the two (HF:739-742)
  XLX_TRY(ensure_device(out));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * L, H = d->hidden;
  // without a save buffer (inference) the pre-LayerNorm sum goes through `out` itself, normalised in place
  EmbSave s = emb_layout(d, M, save);
  float* y = save ? s.y : out;
  XLX_TRY(embed_sum(input_ids, token_type_ids, input_ids ? params[0] : inputs_embeds, params[1], params[2], M, L, H, y, st));
  DropSite drop;
  if (dropout && dropout->p_hidden > 0.f) {
    if (!(dropout->p_hidden < 1.f) || !save) return -1;      // dropout is a training-forward thing
    drop = make_site(dropout->seed, DROP_SITE_EMB, dropout->p_hidden);
  }
  return layernorm_fwd(y, params[3], params[4], d->ln_eps, M, H, 1.0f, nullptr, Split(), out, save ? s.mean : nullptr,
                       save ? s.rstd : nullptr, st, drop);
}

int32_t xlx_embeddings_bwd(const xlx_dims* d, int32_t B, int32_t L, int32_t vocab, int32_t max_pos,
                           int32_t type_vocab, const int64_t* input_ids, const int64_t* token_type_ids,
                           const float* const* params, const void* save, const float* d_out, float* const* grads,
                           float* d_inputs_embeds, void* scratch, size_t scratch_bytes, const xlx_dropout* dropout,
                           void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1 || L < 1 || L > max_pos) return -21;
  if ((!input_ids == !d_inputs_embeds) || !params || !save || !d_out || !grads || !scratch) return -24;
  if (scratch_bytes < xlx_embeddings_scratch_bytes(d, B, L)) return -23;
  XLX_TRY(ensure_device(scratch));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * L, H = d->hidden;
  EmbSave s = emb_layout(d, M, const_cast<void*>(save));
  Bump b; b.base = static_cast<char*>(scratch);
  float* dy = b.f32(static_cast<size_t>(M) * H);
  float* part = b.f32(3 * static_cast<size_t>(reduce_max_blocks()) * H);
  int nblk = 0;
  DropSite drop;
  if (dropout && dropout->p_hidden > 0.f) drop = make_site(dropout->seed, DROP_SITE_EMB, dropout->p_hidden);
  XLX_TRY(layernorm_bwd(d_out, 1.0f, s.y, params[3], s.mean, s.rstd, M, H, dy, Split(), part, &nblk, st, drop));
  float* o[2] = {grads[3], grads[4]};
  XLX_TRY(colsum_finish(part, 2, nblk, H, o, 0, st));
  if (d_inputs_embeds)      // gradient wrt inputs_embeds = the LayerNorm-input gradient itself
    XLX_CUDA(cudaMemcpyAsync(d_inputs_embeds, dy, static_cast<size_t>(M) * H * 4, cudaMemcpyDeviceToDevice, st));
  else
    XLX_CUDA(cudaMemsetAsync(grads[0], 0, static_cast<size_t>(vocab) * H * 4, st));
  XLX_CUDA(cudaMemsetAsync(grads[1], 0, static_cast<size_t>(max_pos) * H * 4, st));
  XLX_CUDA(cudaMemsetAsync(grads[2], 0, static_cast<size_t>(type_vocab) * H * 4, st));
  return embed_scatter(input_ids, token_type_ids, dy, M, L, H, grads[0], grads[1], grads[2], st);
}

// ---- pooler -------------------------------------------------------------------------------------
size_t xlx_pooler_workspace_bytes(const xlx_dims* d, int32_t B) {
  return (hidden_ok(d) && B > 0) ? pool_layout(d, B, nullptr).bytes : 0;
}

int32_t xlx_pooler_fwd(const xlx_dims* d, int32_t B, int32_t L, const float* lang_out, const float* W,
                       const float* bias, float* pooled, void* workspace, size_t workspace_bytes, int32_t passes,
                       void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1 || L < 1) return -21;
  if (!lang_out || !W || !bias || !pooled || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H = d->hidden;
  PoolWs w = pool_layout(d, B, workspace);
  if (w.bytes > workspace_bytes) return -23;
  XLX_TRY(split_rows_f32(lang_out, static_cast<size_t>(L) * H, B, H, w.x0, st));   // hidden_states[:, 0] (HF:577)
  XLX_TRY(split_f32(W, w.w, static_cast<size_t>(H) * H, st));
  GemmEpilogue e;
  e.bias = bias; e.flags = EPI_TANH; e.out_f32 = pooled; e.ld_out = H;
  return gemm_linear(passes, st, w.x0, B, H, w.w, H, e);
}

int32_t xlx_pooler_bwd(const xlx_dims* d, int32_t B, int32_t L, const float* pooled, const float* d_pooled,
                       float* d_lang_out, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                       int32_t passes, void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1 || L < 1) return -21;
  if (!pooled || !d_pooled || !d_lang_out || !dW || !dbias || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H = d->hidden;
  PoolWs w = pool_layout(d, B, workspace);
  if (w.bytes > workspace_bytes) return -23;
  XLX_TRY(tanh_bwd_split(d_pooled, pooled, w.dpre, static_cast<size_t>(B) * H, st));
  // dbias: column sums of dpre; the first H² floats of dW double as the partial-sum scratch before the wgrad
  // GEMM overwrites them (needs reduce_max_blocks-independent bound: colsum uses ≤ 128 row blocks ≤ H rows)
  XLX_TRY(colsum(nullptr, w.dpre, B, H, H, dW, dbias, st));
  XLX_TRY(gemm_wgrad(passes, st, w.dpre, B, H, w.x0, H, dW));
  XLX_CUDA(cudaMemsetAsync(d_lang_out, 0, static_cast<size_t>(B) * L * H * 4, st));
  GemmEpilogue e;
  e.out_f32 = d_lang_out; e.ld_out = L * H;     // row b of the GEMM lands on token 0 of sample b
  return gemm_dgrad(passes, st, w.dpre, B, H, w.w, H, e);
}


// ---- cluster head (lxrt/modeling.py:8-53), LM head (HF:597-607), answer head (HF:610-623) ----------------------
#define XLX_HEAD_API(NAME, KIND)                                                                                     \
  size_t xlx_##NAME##_prep_bytes(const xlx_dims* d, int32_t classes) {                                              \
    HeadDims h;                                                                                                      \
    return head_dims(d, classes, KIND, &h) ? head_prep_layout(h, nullptr).bytes : 0;                                 \
  }                                                                                                                  \
  int32_t xlx_##NAME##_prepare(const xlx_dims* d, int32_t classes, const float* const* params, void* prep,           \
                               void* stream) {                                                                       \
    HeadDims h;                                                                                                      \
    if (!head_dims(d, classes, KIND, &h)) return -20;                                                                \
    if (!params || !prep) return -24;                                                                                \
    XLX_TRY(ensure_device(prep));                                                                                    \
    return head_prepare(h, params, prep, static_cast<cudaStream_t>(stream));                                         \
  }                                                                                                                  \
  size_t xlx_##NAME##_workspace_bytes(const xlx_dims* d, int32_t classes, int32_t M) {                               \
    HeadDims h;                                                                                                      \
    return (head_dims(d, classes, KIND, &h) && M > 0) ? head_ws_layout(h, M, nullptr).bytes : 0;                     \
  }

XLX_HEAD_API(objhead, HEAD_CLUSTER)
XLX_HEAD_API(lmhead, HEAD_LM)
XLX_HEAD_API(qahead, HEAD_ANSWER)

int32_t xlx_objhead_fwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                        const float* hidden, const int64_t* labels, const float* feat_target,
                        const float* feat_weight, float* feat, float* logits, float* loss, float* feat_loss,
                        float* pred_prob, int64_t* pred_id, void* workspace, size_t workspace_bytes, int32_t passes,
                        void* stream) {
  HeadDims h;
  if (!head_dims(d, classes, HEAD_CLUSTER, &h)) return -20;
  if (M < 1) return -21;
  if (!params || !prep || !hidden || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  FeatLoss fl;
  fl.target = feat_target; fl.weight = feat_weight; fl.loss = feat_loss;
  return head_fwd(h, params, prep, M, hidden, labels, fl, feat, logits, loss, pred_prob, pred_id, workspace,
                  workspace_bytes, passes, static_cast<cudaStream_t>(stream));
}
int32_t xlx_objhead_bwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep, int32_t M,
                        const int64_t* labels, const float* d_loss, const float* d_logits, const float* feat_target,
                        const float* feat_weight, const float* d_feat_loss, const float* d_feat, float* d_hidden,
                        float* const* grads, void* workspace, size_t workspace_bytes, int32_t passes, void* stream) {
  HeadDims h;
  if (!head_dims(d, classes, HEAD_CLUSTER, &h)) return -20;
  if (M < 1) return -21;
  if (!params || !prep || !d_hidden || !grads || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  FeatLoss fl;
  fl.target = feat_target; fl.weight = feat_weight; fl.d_loss = d_feat_loss;
  return head_bwd(h, params, prep, M, labels, d_loss, d_logits, fl, d_feat, d_hidden, grads, workspace,
                  workspace_bytes, passes, static_cast<cudaStream_t>(stream));
}
int32_t xlx_feat_row_weight(const uint8_t* vis_mask, int32_t B, int32_t V, int32_t feat_dim, const int64_t* rows,
                            int32_t n, float* weight, void* stream) {
  if (B < 1 || V < 1 || feat_dim < 1 || n < 0 || (!rows && n != B * V)) return -21;
  if (!vis_mask || !weight) return -24;
  XLX_TRY(ensure_device(weight));
  return feat_row_weight(vis_mask, B, V, feat_dim, rows, n, weight, static_cast<cudaStream_t>(stream));
}

#define XLX_PLAIN_HEAD_FWD_BWD(NAME, KIND)                                                                           \
  int32_t xlx_##NAME##_fwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep,         \
                           int32_t M, const float* hidden, const int64_t* labels, float* scores, float* loss,        \
                           float* pred_prob, int64_t* pred_id, void* workspace, size_t workspace_bytes,              \
                           int32_t passes, void* stream) {                                                           \
    HeadDims h;                                                                                                      \
    if (!head_dims(d, classes, KIND, &h)) return -20;                                                                \
    if (M < 1) return -21;                                                                                           \
    if (!params || !prep || !hidden || !workspace || (!pred_prob != !pred_id)) return -24;                           \
    if (!labels && !scores && !pred_prob) return -24;                                                                \
    if (passes != 1 && passes != 3) return -1;                                                                       \
    XLX_TRY(ensure_device(workspace));                                                                               \
    return head_fwd(h, params, prep, M, hidden, labels, FeatLoss(), nullptr, scores, loss, pred_prob, pred_id,       \
                    workspace, workspace_bytes, passes, static_cast<cudaStream_t>(stream));                          \
  }                                                                                                                  \
  int32_t xlx_##NAME##_bwd(const xlx_dims* d, int32_t classes, const float* const* params, const void* prep,         \
                           int32_t M, const int64_t* labels, const float* d_loss, const float* d_logits,             \
                           float* d_hidden, float* const* grads, void* workspace, size_t workspace_bytes,            \
                           int32_t passes, void* stream) {                                                           \
    HeadDims h;                                                                                                      \
    if (!head_dims(d, classes, KIND, &h)) return -20;                                                                \
    if (M < 1) return -21;                                                                                           \
    if (!params || !prep || (!labels && !d_logits) || !d_hidden || !grads || !workspace) return -24;                 \
    if (passes != 1 && passes != 3) return -1;                                                                       \
    XLX_TRY(ensure_device(workspace));                                                                               \
    return head_bwd(h, params, prep, M, labels, d_loss, d_logits, FeatLoss(), nullptr, d_hidden, grads, workspace,   \
                    workspace_bytes, passes, static_cast<cudaStream_t>(stream));                                     \
  }

XLX_PLAIN_HEAD_FWD_BWD(lmhead, HEAD_LM)
XLX_PLAIN_HEAD_FWD_BWD(qahead, HEAD_ANSWER)

// ---- matched head (HF:661,664: seq_relationship = Linear(H, 2) on the pooled output) + its cross-entropy -------
// (lxrt/modeling.py:227-235).  scratch: xlx_matchhead_scratch_floats(B) floats, passed unchanged to the backward.
int64_t xlx_matchhead_scratch_floats(int32_t B) { return B > 0 ? 3 * static_cast<int64_t>(B) + 8 : 0; }

int32_t xlx_matchhead_fwd(const xlx_dims* d, int32_t B, const float* pooled, const float* W, const float* bias,
                          const int64_t* labels, float* scores, float* loss, float* scratch, void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1) return -21;
  if (!pooled || !W || !bias || !scores || (labels && !scratch)) return -24;
  XLX_TRY(ensure_device(scores));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XLX_TRY(small_linear_fwd(pooled, W, bias, B, d->hidden, 2, scores, st));
  if (labels) {
    XLX_TRY(small_ce_fwd(scores, B, 2, labels, -100, scratch + 8, scratch, st));
    if (loss) XLX_CUDA(cudaMemcpyAsync(loss, scratch, 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}
int32_t xlx_matchhead_bwd(const xlx_dims* d, int32_t B, const float* pooled, const float* W, const int64_t* labels,
                          const float* scores, const float* d_loss, float* d_pooled, float* dW, float* dbias,
                          float* scratch, void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1) return -21;
  if (!pooled || !W || !labels || !scores || !d_loss || !d_pooled || !dW || !dbias || !scratch) return -24;
  XLX_TRY(ensure_device(scores));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dscores = scratch + 8 + B;
  XLX_TRY(small_ce_bwd(scores, B, 2, labels, -100, scratch, d_loss, dscores, st));
  return small_linear_bwd(dscores, pooled, W, B, d->hidden, 2, dW, dbias, d_pooled, st);
}

// backward of the bare scores (a caller that brings its own loss): d_scores [B,2] → d_pooled, dW, dbias
int32_t xlx_matchhead_bwd_scores(const xlx_dims* d, int32_t B, const float* pooled, const float* W,
                                 const float* d_scores, float* d_pooled, float* dW, float* dbias, void* stream) {
  if (!hidden_ok(d)) return -20;
  if (B < 1) return -21;
  if (!pooled || !W || !d_scores || !d_pooled || !dW || !dbias) return -24;
  XLX_TRY(ensure_device(d_pooled));
  return small_linear_bwd(d_scores, pooled, W, B, d->hidden, 2, dW, dbias, d_pooled, static_cast<cudaStream_t>(stream));
}

// ---- row compaction around the masked-prediction losses --------------------------------------------------------
// CrossEntropyLoss skips rows labelled −100 (lxrt/modeling.py:99,102,253-256); 50 % of the visual rows and 85 % of the
// word rows are.  Compacting before the head keeps the loss and every gradient identical and skips their GEMM work.
int32_t xlx_labelled_rows(const int64_t* labels, int32_t M, int64_t ignore_index, int64_t* rows, int32_t* count,
                          void* stream) {
  if (M < 1) return -21;
  if (!labels || !rows || !count) return -24;
  XLX_TRY(ensure_device(rows));
  return labelled_rows(labels, M, ignore_index, rows, count, static_cast<cudaStream_t>(stream));
}
int32_t xlx_gather_rows(const float* src, const int64_t* labels, const int64_t* rows, int32_t n, int32_t cols,
                        float* dst, int64_t* labels_dst, void* stream) {
  if (n < 0 || (src && (cols < 4 || cols % 4))) return -21;
  if (!rows || (!src != !dst) || (!labels != !labels_dst) || (!src && !labels)) return -24;
  if (n == 0) return 0;
  XLX_TRY(ensure_device(rows));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src) XLX_TRY(gather_rows(src, rows, nullptr, nullptr, n, cols, dst, Split{}, st));
  if (labels) XLX_TRY(gather_i64(labels, rows, n, labels_dst, st));
  return 0;
}
int32_t xlx_scatter_rows(const float* src, const int64_t* rows, int32_t n, int32_t M, int32_t cols, float* dst,
                         void* stream) {
  if (n < 0 || M < 1 || n > M || cols < 4 || cols % 4) return -21;
  if (!src || !rows || !dst) return -24;
  XLX_TRY(ensure_device(dst));
  return scatter_rows(src, rows, n, M, cols, dst, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
