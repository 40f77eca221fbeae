// Host-side orchestration of the LXMERT encoder forward / backward over the sm_100a kernels, and the
// C ABI declared in include/xlxmert_b200.h.  Follows HF modeling_lxmert.py:487-565 (LxmertEncoder),
// :353-366 (LxmertLayer), :369-457 (LxmertXLayer), :460-484 (LxmertVisualFeatureEncoder) — the arithmetic
// the reference reaches through self.bert at x-lxmert/src/lxrt/modeling.py:195-206.
#include <cstring>
#include <vector>

#include "../../include/xlxmert_b200.h"
#include "host_util.cuh"

using namespace xlx;

namespace {

constexpr int N_ATT = 10, N_FFN = 6, N_VISN = 8;

bool dims_ok(const xlx_dims* d) {
  return d && d->hidden > 0 && d->hidden % 128 == 0 && d->hidden <= 1024 && d->heads * 64 == d->hidden &&
         d->intermediate % 8 == 0 && d->feat_dim % 8 == 0 && d->pos_dim == 4 && d->l_layers >= 0 &&
         d->r_layers >= 0 && d->x_layers >= 0;
}

// ---- parameter slots --------------------------------------------------------------------------
struct SlotTable {
  std::vector<int64_t> elems, offset;
  int64_t total = 0;
};
void push_att(std::vector<int64_t>& e, int64_t H) {
  const int64_t s[N_ATT] = {H * H, H * H, H * H, H, H, H, H * H, H, H, H};
  e.insert(e.end(), s, s + N_ATT);
}
void push_ffn(std::vector<int64_t>& e, int64_t H, int64_t I) {
  const int64_t s[N_FFN] = {I * H, I, H * I, H, H, H};
  e.insert(e.end(), s, s + N_FFN);
}
SlotTable slot_table(const xlx_dims* d) {
  SlotTable t;
  const int64_t H = d->hidden, I = d->intermediate, F = d->feat_dim;
  const int64_t v[N_VISN] = {H * F, H, H, H, H * 4, H, H, H};
  t.elems.insert(t.elems.end(), v, v + N_VISN);
  for (int i = 0; i < d->l_layers + d->r_layers; ++i) { push_att(t.elems, H); push_ffn(t.elems, H, I); }
  for (int i = 0; i < d->x_layers; ++i) {
    push_att(t.elems, H); push_att(t.elems, H); push_att(t.elems, H);
    push_ffn(t.elems, H, I); push_ffn(t.elems, H, I);
  }
  t.offset.resize(t.elems.size());
  for (size_t i = 0; i < t.elems.size(); ++i) { t.offset[i] = t.total; t.total += t.elems[i]; }
  return t;
}
int att_slot(const xlx_dims* d, int blk) {  // slot index of attention block `blk` (plan order)
  const int nlr = d->l_layers + d->r_layers;
  if (blk < nlr) return N_VISN + blk * (N_ATT + N_FFN);
  const int k = (blk - nlr) / 3, w = (blk - nlr) % 3;
  return N_VISN + nlr * (N_ATT + N_FFN) + k * (3 * N_ATT + 2 * N_FFN) + w * N_ATT;
}
int ffn_slot(const xlx_dims* d, int blk) {
  const int nlr = d->l_layers + d->r_layers;
  if (blk < nlr) return N_VISN + blk * (N_ATT + N_FFN) + N_ATT;
  const int k = (blk - nlr) / 2, w = (blk - nlr) % 2;
  return N_VISN + nlr * (N_ATT + N_FFN) + k * (3 * N_ATT + 2 * N_FFN) + 3 * N_ATT + w * N_FFN;
}

// ---- prepared weights ----------------------------------------------------------------------------
// Every Linear weight W [N, K] is kept twice as split bf16: as is (forward: Y = X·Wᵀ reads it K-major) and
// transposed Wᵀ [K, N] (dgrad: dX = dY·W reads Wᵀ K-major too) — K-major B operands are ≈ 20 % faster than
// MN-major ones on this kernel, and one batched launch per step builds both.
struct AttW { Split wqkv, wqkv_t; float* bqkv; Split wo, wo_t; };
struct FfnW { Split w1, w1_t, w2, w2_t; };
struct Prep {
  Split visn_w, visn_w_t;
  std::vector<AttW> att;
  std::vector<FfnW> ffn;
  int max_jobs = 0;
  size_t bytes = 0;
};
Prep prep_layout(const xlx_dims* d, void* base) {
  Prep p;
  Bump b;
  b.base = static_cast<char*>(base);
  const size_t H = d->hidden, I = d->intermediate, F = d->feat_dim;
  p.visn_w = b.split(H * F);
  p.visn_w_t = b.split(F * H);
  const int natt = d->l_layers + d->r_layers + 3 * d->x_layers;
  const int nffn = d->l_layers + d->r_layers + 2 * d->x_layers;
  p.att.resize(natt);
  p.ffn.resize(nffn);
  for (auto& a : p.att) {
    a.wqkv = b.split(3 * H * H); a.wqkv_t = b.split(3 * H * H); a.bqkv = b.f32(3 * H);
    a.wo = b.split(H * H); a.wo_t = b.split(H * H);
  }
  for (auto& f : p.ffn) { f.w1 = b.split(I * H); f.w1_t = b.split(H * I); f.w2 = b.split(H * I); f.w2_t = b.split(I * H); }
  p.max_jobs = 1 + 4 * natt + 2 * nffn;
  p.bytes = b.off + 256;
  return p;
}

// ---- activation plan -----------------------------------------------------------------------------
struct AttSave {
  Split qkv;               // [M, 3H] fused Q | K | V projections, split bf16 (the attention kernels TMA-load it)
  float* probs = nullptr;  // self: [B,heads,S,S]; cross: lang-query probs [B,heads,L,V]
  float* probs2 = nullptr; // cross only: vis-query probs [B,heads,V,L]
  Split ctx;               // [M, H]
  float* y = nullptr;      // pre-LN [M, H]
  float* mean = nullptr;
  float* rstd = nullptr;
  Split in, out;           // block input / LN output (row-sliced views)
  int M = 0;
};
struct FfnSave {
  float* u = nullptr;  // gelu'(pre-activation) [M, I], saved by the forward epilogue for the backward (bf16 storage was
                       // tried: no measurable speed-up and the embedding gradient left the 1e-3 tolerance)
  Split h;             // gelu(u) [M, I]
  float* y = nullptr;
  float* mean = nullptr;
  float* rstd = nullptr;
  Split in, out;
  int M = 0;
};
struct BwdScratch {
  Split dy_s;                // LayerNorm-input gradient [M, H]: GEMM operand and residual-path addend
  Split dy_m;                // the same gradient ∘ dropout mask of the dense output (training with dropout): GEMM operand
  Split dqkv, du, dctx;
  float* part = nullptr;     // partial column sums (bias / LayerNorm-affine gradients)
  float* splitk = nullptr;   // split-K partial sums of the weight-gradient GEMMs
};
struct Plan {
  int B, L, V, H, I, F, heads, Ml, Mv, Mt;
  bool training;
  Split feats, lang0, vis0;
  float *y1 = nullptr, *y2 = nullptr, *tbox = nullptr, *vstats = nullptr;
  std::vector<AttSave> att;
  std::vector<FfnSave> ffn;
  size_t part_elems = 0;
  // backward scratch: set 0 (all B·(L+V) rows) serves the caller's stream, set 1 (language rows only) the side stream
  float *dA = nullptr, *dB = nullptr, *dy2 = nullptr;
  float* part_arena = nullptr;     // deferred partial column sums of one backward call (see Deferred)
  size_t part_arena_floats = 0;
  BwdScratch sc[2];
  size_t bytes = 0;
};

Plan make_plan(const xlx_dims* d, int B, int L, int V, bool training, void* base) {
  Plan p;
  p.B = B; p.L = L; p.V = V; p.H = d->hidden; p.I = d->intermediate; p.F = d->feat_dim; p.heads = d->heads;
  p.Ml = B * L; p.Mv = B * V; p.Mt = p.Ml + p.Mv; p.training = training;
  const size_t H = p.H, I = p.I, F = p.F, Ml = p.Ml, Mv = p.Mv, Mt = p.Mt;
  const size_t Mmax = Ml > Mv ? Ml : Mv;
  Bump b;
  b.base = static_cast<char*>(base);
  const int nl = d->l_layers, nr = d->r_layers, nx = d->x_layers;
  p.att.resize(nl + nr + 3 * nx);
  p.ffn.resize(nl + nr + 2 * nx);

  size_t pe = 3 * static_cast<size_t>(reduce_max_blocks()) * H;
  size_t widest = 3 * H > I ? 3 * H : I;
  if (F > widest) widest = F;
  if (128 * widest > pe) pe = 128 * widest;
  if (5 * 129 * H > pe) pe = 5 * 129 * H;
  const size_t row_groups = 4 * ((Mmax + 127) / 128);          // GEMM-epilogue column-sum partials: one row per 32 rows
  if (row_groups * widest > pe) pe = row_groups * widest;
  p.part_elems = pe;
  p.sc[0].part = b.f32(pe);

  p.feats = b.split(Mv * F);
  p.y1 = b.f32(Mv * H);
  p.y2 = b.f32(Mv * H);
  p.tbox = b.f32(Mv * H);
  p.vstats = b.f32(4 * Mv);
  p.lang0 = b.split(Ml * H);
  p.vis0 = b.split(Mv * H);
  Split xin0 = b.split(Mt * H);  // input of the first cross-modality layer: [lang rows | vis rows]

  // inference: the blocks of one modality share one set of temporaries and layer outputs rotate through a small ring.
  // Language blocks use rows [0, Ml) and vision blocks rows [Ml, Mt) of every shared buffer, with their own ring
  // positions, because the two chains run concurrently on two streams (see SideStream below).
  float* sh_y = nullptr;
  Split sh_qkv, sh_ctx, sh_h[2], ring[4];
  int ring_i[3] = {0, 0, 0};        // language / vision / joint
  if (!training) {
    sh_qkv = b.split(Mt * 3 * H);
    sh_ctx = b.split(Mt * H);
    sh_y = b.f32(Mt * H);
    sh_h[0] = b.split(Ml * I);
    sh_h[1] = b.split(Mv * I);
    for (auto& r : ring) r = b.split(Mt * H);
  }
  // who: 0 = language rows, 1 = vision rows, 2 = all rows
  auto new_state = [&](size_t M, int who) -> Split {
    if (training) return b.split(M * H);
    Split s = ring[ring_i[who]];
    ring_i[who] = (ring_i[who] + 1) & 3;
    return who == 1 ? rows(s, Ml, H) : s;
  };
  auto fill_att = [&](AttSave& a, size_t M, size_t probs_elems, size_t probs2_elems, int who) {
    a.M = static_cast<int>(M);
    if (training) {
      a.qkv = b.split(M * 3 * H);
      a.probs = b.f32(probs_elems);
      if (probs2_elems) a.probs2 = b.f32(probs2_elems);
      a.ctx = b.split(M * H);
      a.y = b.f32(M * H);
      a.mean = b.f32(M);
      a.rstd = b.f32(M);
    } else {
      const size_t row0 = who == 1 ? Ml : 0;
      a.qkv = rows(sh_qkv, row0, 3 * H); a.ctx = rows(sh_ctx, row0, H); a.y = sh_y + row0 * H;
    }
  };
  auto fill_ffn = [&](FfnSave& f, size_t M, int who) {
    f.M = static_cast<int>(M);
    if (training) {
      f.u = b.f32(M * I);
      f.h = b.split(M * I);
      f.y = b.f32(M * H);
      f.mean = b.f32(M);
      f.rstd = b.f32(M);
    } else {
      f.h = sh_h[who]; f.y = sh_y + (who == 1 ? Ml : 0) * H;
    }
  };
  const size_t nh = p.heads;
  // single-modality stacks; the last layer of each writes its output straight into xin0
  for (int s = 0; s < 2; ++s) {
    const int n = s ? nr : nl, blk0 = s ? nl : 0;
    const size_t M = s ? Mv : Ml, S = s ? V : L, row0 = s ? Ml : 0;
    Split cur = s ? p.vis0 : p.lang0;
    if (n == 0) {  // no layers of this kind: the embedding output itself is the cross-layer input
      if (s) p.vis0 = rows(xin0, row0, H); else p.lang0 = rows(xin0, row0, H);
    }
    for (int i = 0; i < n; ++i) {
      AttSave& a = p.att[blk0 + i];
      FfnSave& f = p.ffn[blk0 + i];
      fill_att(a, M, B * nh * S * S, 0, s);
      a.in = cur;
      a.out = new_state(M, s);
      fill_ffn(f, M, s);
      f.in = a.out;
      f.out = (i == n - 1) ? rows(xin0, row0, H) : new_state(M, s);
      cur = f.out;
    }
  }
  Split X = xin0;
  for (int k = 0; k < nx; ++k) {
    AttSave& c = p.att[nl + nr + 3 * k];
    AttSave& sl = p.att[nl + nr + 3 * k + 1];
    AttSave& sv = p.att[nl + nr + 3 * k + 2];
    FfnSave& fl = p.ffn[nl + nr + 2 * k];
    FfnSave& fv = p.ffn[nl + nr + 2 * k + 1];
    fill_att(c, Mt, B * nh * L * V, B * nh * V * L, 2);
    c.in = X;
    c.out = new_state(Mt, 2);
    Split S2 = new_state(Mt, 2);
    fill_att(sl, Ml, B * nh * L * L, 0, 0);
    fill_att(sv, Mv, B * nh * V * V, 0, 1);
    sl.in = rows(c.out, 0, H); sl.out = rows(S2, 0, H);
    sv.in = rows(c.out, Ml, H); sv.out = rows(S2, Ml, H);
    Split O = new_state(Mt, 2);
    fill_ffn(fl, Ml, 0);
    fill_ffn(fv, Mv, 1);
    fl.in = sl.out; fl.out = rows(O, 0, H);
    fv.in = sv.out; fv.out = rows(O, Ml, H);
    X = O;
  }
  if (training) {
    p.dA = b.f32(Mt * H);
    p.dB = b.f32(Mt * H);
    p.dy2 = b.f32(Mv * H);
    {
      // every block's partials of a whole backward: LayerNorm (3 vectors × reduce_max_blocks rows), intermediate bias
      // (one row per 32 rows of the block), Q/K/V bias (≤ 128 rows of 3H)
      auto r64 = [](size_t n) { return (n + 63) & ~static_cast<size_t>(63); };
      size_t need = 0;
      const size_t ln = r64(3 * static_cast<size_t>(reduce_max_blocks()) * H), qb = r64(128 * 3 * H);
      auto ffn_rows = [&](size_t M) { return r64(4 * ((M + 127) / 128) * I); };
      need += static_cast<size_t>(nl + nx) * (2 * ln + qb + ffn_rows(Ml));         // language-side blocks
      need += static_cast<size_t>(nr + nx) * (2 * ln + qb + ffn_rows(Mv));         // vision-side blocks
      need += static_cast<size_t>(nx) * (ln + qb);                                 // cross-attention blocks
      p.part_arena_floats = need;
      p.part_arena = b.f32(need);
    }
    for (int i = 0; i < 2; ++i) {
      BwdScratch& c = p.sc[i];
      const size_t M = i ? Ml : Mt;
      c.dctx = b.split(M * H);
      c.dy_s = b.split(M * H);
      c.dy_m = b.split(M * H);
      c.dqkv = b.split(M * 3 * H);
      c.du = b.split((i ? Ml : Mmax) * I);
      c.splitk = b.f32(gemm_splitk_ws_floats());
      if (i) c.part = b.f32(pe);
    }
  }
  p.bytes = b.off + 256;
  return p;
}

// ---- GEMM wrappers -------------------------------------------------------------------------------
struct Run {
  const xlx_dims* d;
  const float* const* params;
  Prep prep;
  Plan plan;
  int passes;
  cudaStream_t st;
  const float* lmask;
  const float* vmask;
  DropoutCfg drop;     // all-zero unless a training forward / its backward was given an xlx_dropout
  int n_att = 0;       // attention blocks in the plan (site numbering)
  DropSite hidden_site(uint32_t site) const { return make_site(drop.seed, site, drop.p_hidden); }
  DropSite probs_site(int blk, int dir) const { return make_site(drop.seed, site_probs(blk, dir), drop.p_attn); }
};

int linear(const Run& r, Split x, int M, int K, Split w, int N, const GemmEpilogue& e) {
  return gemm_linear(r.passes, r.st, x, M, K, w, N, e);
}
// dX[M,K] = dY[M,N] · W[N,K], with the transposed copy Wᵀ [K, N] as a K-major B operand
int dgrad(const Run& r, Split dy, int M, int N, Split w_t, int K, const GemmEpilogue& e) {
  return gemm_linear(r.passes, r.st, dy, M, N, w_t, K, e);
}
int wgrad(const Run& r, float* splitk, Split dy, int M, int N, Split x, int K, float* dw) {
  return gemm_wgrad(r.passes, r.st, dy, M, N, x, K, dw, false, 0, splitk);
}

const float* P(const Run& r, int slot) { return r.params[slot]; }

// attention operands: part 0/1/2 = Q/K/V columns of a fused [M, 3H] projection; or a plain [M, H] matrix
AttnOperand qkv_op(Split qkv, int M, int H, int part) {
  AttnOperand o;
  o.base = qkv; o.ld = 3 * H; o.rows = M; o.col = part * H;
  return o;
}
AttnOperand mat_op(Split m, int M, int H) {
  AttnOperand o;
  o.base = m; o.ld = H; o.rows = M; o.col = 0;
  return o;
}

// ---- forward blocks ------------------------------------------------------------------------------
// residual + LayerNorm tail shared by the attention-output and FFN-output sub-blocks (HF:277-288, 339-350)
int ln_tail(const Run& r, float* y, int M, const float* g, const float* b, Split out, float* out_f32, float* mean,
            float* rstd) {
  return layernorm_fwd(y, g, b, r.d->ln_eps, M, r.plan.H, 1.0f, nullptr, out, out_f32, mean, rstd, r.st);
}

int att_self_fwd(const Run& r, int blk, int S, const float* mask, float* out_f32) {
  const Plan& p = r.plan;
  const AttSave& a = p.att[blk];
  const AttW& w = r.prep.att[blk];
  const int s0 = att_slot(r.d, blk), H = p.H, M = a.M;
  GemmEpilogue e;
  e.bias = w.bqkv; e.out_hi = a.qkv.hi; e.out_lo = a.qkv.lo; e.ld_split = 3 * H;
  XLX_TRY(linear(r, a.in, M, H, w.wqkv, 3 * H, e));
  XLX_TRY(attention_fwd(qkv_op(a.qkv, M, H, 0), qkv_op(a.qkv, M, H, 1), qkv_op(a.qkv, M, H, 2), mask, p.B, p.heads, S, S,
                        a.ctx, nullptr, H, a.probs, r.st, r.probs_site(blk, 0)));
  GemmEpilogue o;
  o.bias = P(r, s0 + 7); o.addend_hi = a.in.hi; o.addend_lo = a.in.lo; o.ld_addend = H;
  o.out_f32 = a.y; o.ld_out = H;
  o.drop = r.hidden_site(site_att_out(blk));
  XLX_TRY(linear(r, a.ctx, M, H, w.wo, H, o));
  return ln_tail(r, a.y, M, P(r, s0 + 8), P(r, s0 + 9), a.out, out_f32, a.mean, a.rstd);
}

// LxmertXLayer.cross_att (HF:385-406): one shared attention module, both directions read the layer inputs.
int att_cross_fwd(const Run& r, int blk) {
  const Plan& p = r.plan;
  const AttSave& a = p.att[blk];
  const AttW& w = r.prep.att[blk];
  const int s0 = att_slot(r.d, blk), H = p.H;
  const size_t H3 = 3 * static_cast<size_t>(H);
  GemmEpilogue e;
  e.bias = w.bqkv; e.out_hi = a.qkv.hi; e.out_lo = a.qkv.lo; e.ld_split = 3 * H;
  XLX_TRY(linear(r, a.in, p.Mt, H, w.wqkv, 3 * H, e));   // Q, K, V of all B·(L+V) tokens in one GEMM
  const Split ql = a.qkv, qv = rows(a.qkv, p.Ml, H3);     // language rows / vision rows
  // language queries over vision keys/values (mask = visual attention mask, normally none)
  XLX_TRY(attention_fwd(qkv_op(ql, p.Ml, H, 0), qkv_op(qv, p.Mv, H, 1), qkv_op(qv, p.Mv, H, 2), r.vmask, p.B, p.heads,
                        p.L, p.V, a.ctx, nullptr, H, a.probs, r.st, r.probs_site(blk, 0)));
  // vision queries over language keys/values (mask = language attention mask)
  XLX_TRY(attention_fwd(qkv_op(qv, p.Mv, H, 0), qkv_op(ql, p.Ml, H, 1), qkv_op(ql, p.Ml, H, 2), r.lmask, p.B, p.heads,
                        p.V, p.L, rows(a.ctx, p.Ml, H), nullptr, H, a.probs2, r.st, r.probs_site(blk, 1)));
  GemmEpilogue o;
  o.bias = P(r, s0 + 7); o.addend_hi = a.in.hi; o.addend_lo = a.in.lo; o.ld_addend = H;
  o.out_f32 = a.y; o.ld_out = H;
  o.drop = r.hidden_site(site_att_out(blk));
  XLX_TRY(linear(r, a.ctx, p.Mt, H, w.wo, H, o));
  return ln_tail(r, a.y, p.Mt, P(r, s0 + 8), P(r, s0 + 9), a.out, nullptr, a.mean, a.rstd);
}

int ffn_fwd(const Run& r, int blk, float* out_f32) {
  const Plan& p = r.plan;
  const FfnSave& f = p.ffn[blk];
  const FfnW& w = r.prep.ffn[blk];
  const int s0 = ffn_slot(r.d, blk), H = p.H, I = p.I, M = f.M;
  GemmEpilogue e;
  e.bias = P(r, s0 + 1); e.flags = EPI_GELU | (f.u ? EPI_SAVE_DGELU : 0); e.out_u = f.u; e.ld_u = I;   // f.u ← gelu'(pre-activation)
  e.out_hi = f.h.hi; e.out_lo = f.h.lo; e.ld_split = I;
  XLX_TRY(linear(r, f.in, M, H, w.w1, I, e));
  GemmEpilogue o;
  o.bias = P(r, s0 + 3); o.addend_hi = f.in.hi; o.addend_lo = f.in.lo; o.ld_addend = H;
  o.out_f32 = f.y; o.ld_out = H;
  o.drop = r.hidden_site(site_ffn_out(r.n_att, blk));
  XLX_TRY(linear(r, f.h, M, I, w.w2, H, o));
  return ln_tail(r, f.y, M, P(r, s0 + 4), P(r, s0 + 5), f.out, out_f32, f.mean, f.rstd);
}

// ---- backward blocks -----------------------------------------------------------------------------
// All take the gradient wrt the block output in `dout` ([M,H] fp32) and leave the gradient wrt the block
// input in `din`.  Parameter gradients go to the flat arena `grads` at the slot offsets.
// Partial column sums (bias / LayerNorm-affine gradients) are not finished block by block: each producer writes its
// partials into its own piece of the workspace's partial arena and registers a FinishJob; one batched launch at the end
// of the backward call finishes them all (colsum_finish_batched) — ≈ 140 launches per step less.
struct Deferred {
  float* arena = nullptr;
  size_t cap = 0, used = 0;
  std::vector<FinishJob> jobs[2];      // [0] registered by blocks on the caller's stream, [1] by language-side blocks
  float* take(size_t n) {                      // nullptr when the arena is exhausted (callers then finish right away)
    n = (n + 63) & ~static_cast<size_t>(63);
    if (used + n > cap) return nullptr;
    float* p = arena + used;
    used += n;
    return p;
  }
};
struct Bwd {
  const Run* r;            // carries the stream the blocks are issued on
  const BwdScratch* sc;    // temporaries owned by that stream
  float* grads;
  SlotTable slots;
  Deferred* def = nullptr; // shared by the two streams' views (host-side bookkeeping only)
  int lane = 0;            // which of the Deferred job lists this view feeds
  float* G(int slot) const { return grads + slots.offset[slot]; }
};
// Finish `part` [nvec, nblk, H] into outs now, or later with the batch when the partials live in the deferred arena.
int finish_or_defer(const Bwd& bw, const float* part, bool in_arena, int nvec, int nblk, int H, float* const* outs) {
  if (in_arena) {
    FinishJob j{};
    j.part = part; j.nvec = nvec; j.nblk = nblk; j.H = H;
    for (int v = 0; v < nvec; ++v) j.out[v] = outs[v];
    bw.def->jobs[bw.lane].push_back(j);
    return 0;
  }
  return colsum_finish(part, nvec, nblk, H, outs, 0, bw.r->st);
}

// LayerNorm backward of a "dense → +residual → LayerNorm" tail; also yields the dense bias gradient (slot_g − 1),
// which is the column sum of the LayerNorm-input gradient.
// With dropout on the dense output (`site`), dy_s keeps the plain gradient (residual path) and bw.sc->dy_m receives
// dy ∘ mask — the operand of the dense layer's weight / input gradients; *dense_dy tells the caller which one to use.
int ln_tail_bwd(const Bwd& bw, const float* dout, const float* y, int slot_g, const float* mean, const float* rstd,
                int M, float* dy, Split dy_s, uint32_t site, Split* dense_dy) {
  const Run& r = *bw.r;
  int nblk = 0;
  const DropSite ds = r.hidden_site(site);
  *dense_dy = ds.threshold ? bw.sc->dy_m : dy_s;
  float* part = bw.def ? bw.def->take(3 * static_cast<size_t>(reduce_max_blocks()) * r.plan.H) : nullptr;
  const bool deferred = part != nullptr;
  if (!deferred) part = bw.sc->part;
  XLX_TRY(layernorm_bwd(dout, 1.0f, y, P(r, slot_g), mean, rstd, M, r.plan.H, dy, dy_s, part, &nblk, r.st,
                        DropSite(), ds, bw.sc->dy_m));
  float* outs[3] = {bw.G(slot_g), bw.G(slot_g + 1), bw.G(slot_g - 1)};
  return finish_or_defer(bw, part, deferred, 3, nblk, r.plan.H, outs);
}

int ffn_bwd(const Bwd& bw, int blk, const float* dout, float* din) {
  const Run& r = *bw.r;
  const Plan& p = r.plan;
  const BwdScratch& c = *bw.sc;
  const FfnSave& f = p.ffn[blk];
  const FfnW& w = r.prep.ffn[blk];
  const int s0 = ffn_slot(r.d, blk), H = p.H, I = p.I, M = f.M;
  Split dyd;    // gradient wrt the dense output (= dy, or dy ∘ dropout mask)
  XLX_TRY(ln_tail_bwd(bw, dout, f.y, s0 + 4, f.mean, f.rstd, M, nullptr, c.dy_s, site_ffn_out(r.n_att, blk), &dyd));  // + bias grad (s0 + 3)
  XLX_TRY(wgrad(r, c.splitk, dyd, M, H, f.h, I, bw.G(s0 + 2)));
  GemmEpilogue e;   // du = (dy · W2) ∘ gelu'(u), the derivative was saved by the forward
  e.flags = EPI_MUL; e.u_in = f.u; e.ld_u = I; e.out_hi = c.du.hi; e.out_lo = c.du.lo; e.ld_split = I;
  // the intermediate bias gradient = column sums of du, gathered by the same epilogue (one partial row per 32 rows)
  float* bpart = bw.def ? bw.def->take(static_cast<size_t>(4 * ((M + 127) / 128)) * I) : nullptr;
  const bool bdef = bpart != nullptr;
  if (!bdef) bpart = c.part;
  e.colsum_part = bpart;
  XLX_TRY(dgrad(r, dyd, M, H, w.w2_t, I, e));
  {
    float* outs[1] = {bw.G(s0 + 1)};
    XLX_TRY(finish_or_defer(bw, bpart, bdef, 1, (M + 31) / 32, I, outs));
  }
  XLX_TRY(wgrad(r, c.splitk, c.du, M, I, f.in, H, bw.G(s0)));
  GemmEpilogue o;   // din = du · W1 + dy (residual path)
  o.addend_hi = c.dy_s.hi; o.addend_lo = c.dy_s.lo; o.ld_addend = H;   // residual path: + dy (kept as split bf16 only)
  o.out_f32 = din; o.ld_out = H;
  return dgrad(r, c.du, M, I, w.w1_t, H, o);
}

// common tail/head of the attention backward: everything except the attention-core call(s)
int att_bwd_head(const Bwd& bw, int blk, const float* dout) {
  const Run& r = *bw.r;
  const Plan& p = r.plan;
  const BwdScratch& c = *bw.sc;
  const AttSave& a = p.att[blk];
  const AttW& w = r.prep.att[blk];
  const int s0 = att_slot(r.d, blk), H = p.H, M = a.M;
  Split dyd;
  XLX_TRY(ln_tail_bwd(bw, dout, a.y, s0 + 8, a.mean, a.rstd, M, nullptr, c.dy_s, site_att_out(blk), &dyd));   // + bias grad (s0 + 7)
  XLX_TRY(wgrad(r, c.splitk, dyd, M, H, a.ctx, H, bw.G(s0 + 6)));
  GemmEpilogue e;
  e.out_hi = c.dctx.hi; e.out_lo = c.dctx.lo; e.ld_split = H;
  return dgrad(r, dyd, M, H, w.wo_t, H, e);
}
int att_bwd_tail(const Bwd& bw, int blk, float* din) {
  const Run& r = *bw.r;
  const Plan& p = r.plan;
  const BwdScratch& c = *bw.sc;
  const AttSave& a = p.att[blk];
  const AttW& w = r.prep.att[blk];
  const int s0 = att_slot(r.d, blk), H = p.H, M = a.M;
  {                                                                              // q.bias | k.bias | v.bias
    float* qpart = bw.def ? bw.def->take(128 * static_cast<size_t>(3 * H)) : nullptr;
    const bool qdef = qpart != nullptr;
    if (!qdef) qpart = c.part;
    int nb = 0;
    XLX_TRY(colsum_partial(nullptr, c.dqkv, M, 3 * H, 3 * H, qpart, &nb, r.st));
    float* outs[1] = {bw.G(s0 + 3)};
    XLX_TRY(finish_or_defer(bw, qpart, qdef, 1, nb, 3 * H, outs));
  }
  XLX_TRY(wgrad(r, c.splitk, c.dqkv, M, 3 * H, a.in, H, bw.G(s0)));                          // q.w | k.w | v.w
  GemmEpilogue o;
  o.addend_hi = c.dy_s.hi; o.addend_lo = c.dy_s.lo; o.ld_addend = H;   // residual path: + dy (kept as split bf16 only)
  o.out_f32 = din; o.ld_out = H;
  return dgrad(r, c.dqkv, M, 3 * H, w.wqkv_t, H, o);
}
int att_self_bwd(const Bwd& bw, int blk, int S, const float* dout, float* din) {
  const Run& r = *bw.r;
  const Plan& p = r.plan;
  const BwdScratch& c = *bw.sc;
  const AttSave& a = p.att[blk];
  const int H = p.H;
  XLX_TRY(att_bwd_head(bw, blk, dout));
  Split dq = c.dqkv, dk = c.dqkv, dv = c.dqkv;
  dk.hi += H; dk.lo += H; dv.hi += 2 * H; dv.lo += 2 * H;
  XLX_TRY(attention_bwd(mat_op(c.dctx, a.M, H), qkv_op(a.qkv, a.M, H, 0), qkv_op(a.qkv, a.M, H, 1),
                        qkv_op(a.qkv, a.M, H, 2), a.probs, p.B, p.heads, S, S, dq, dk, dv, 3 * H, r.st, r.probs_site(blk, 0)));
  return att_bwd_tail(bw, blk, din);
}
int att_cross_bwd(const Bwd& bw, int blk, const float* dout, float* din) {
  const Run& r = *bw.r;
  const Plan& p = r.plan;
  const BwdScratch& c = *bw.sc;
  const AttSave& a = p.att[blk];
  const int H = p.H;
  const size_t H3 = 3 * static_cast<size_t>(H);
  XLX_TRY(att_bwd_head(bw, blk, dout));
  const Split ql = a.qkv, qv = rows(a.qkv, p.Ml, H3);
  Split dl = c.dqkv, dvv = rows(c.dqkv, p.Ml, H3);   // language rows / vision rows of dqkv
  auto col = [](Split s, int c) { s.hi += c; s.lo += c; return s; };
  // language queries: dQ → language rows, dK/dV → vision rows
  XLX_TRY(attention_bwd(mat_op(c.dctx, p.Ml, H), qkv_op(ql, p.Ml, H, 0), qkv_op(qv, p.Mv, H, 1), qkv_op(qv, p.Mv, H, 2),
                        a.probs, p.B, p.heads, p.L, p.V, col(dl, 0), col(dvv, H), col(dvv, 2 * H), 3 * H, r.st,
                        r.probs_site(blk, 0)));
  // vision queries: dQ → vision rows, dK/dV → language rows
  XLX_TRY(attention_bwd(mat_op(rows(c.dctx, p.Ml, H), p.Mv, H), qkv_op(qv, p.Mv, H, 0), qkv_op(ql, p.Ml, H, 1),
                        qkv_op(ql, p.Ml, H, 2), a.probs2, p.B, p.heads, p.V, p.L, col(dvv, 0), col(dl, H),
                        col(dl, 2 * H), 3 * H, r.st, r.probs_site(blk, 1)));
  return att_bwd_tail(bw, blk, din);
}

// ---- second stream -------------------------------------------------------------------------------
// The language and the vision stacks are independent between the cross-attention blocks.  Language GEMMs have
// M = B·L = 5120 rows at the reference batch: 120 or 480 tiles on 148 SMs, i.e. 19 % of every such launch is idle tail.
// Issuing the language-side blocks on a library-owned side stream lets the two chains fill each other's tails.  The
// caller's stream stays the only externally visible one: fork = side waits for an event on it, join = it waits for the
// side stream (the pattern is CUDA-graph capturable).  XLX_TWO_STREAMS=0 issues everything on the caller's stream.
struct SideStream {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
  static const bool on = [] { const char* e = getenv("XLX_TWO_STREAMS"); return !(e && e[0] == '0'); }();
  if (!on || gemm_timing_active()) return nullptr;   // per-launch timing wants non-overlapping kernels
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (!s.st) {
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      s.st = nullptr;
      cudaGetLastError();
      return nullptr;
    }
  }
  return &s;
}
// side stream continues from the current end of `main`
int fork_to(SideStream* s, cudaStream_t main) {
  if (!s) return 0;
  XLX_CUDA(cudaEventRecord(s->fork, main));
  XLX_CUDA(cudaStreamWaitEvent(s->st, s->fork, 0));
  return 0;
}
// `main` continues after everything issued on the side stream so far
int join_from(SideStream* s, cudaStream_t main) {
  if (!s) return 0;
  XLX_CUDA(cudaEventRecord(s->join, s->st));
  XLX_CUDA(cudaStreamWaitEvent(main, s->join, 0));
  return 0;
}

int check_common(const xlx_dims* d, int B, int L, int V) {
  if (!dims_ok(d)) return -20;
  if (B < 1 || L < 1 || V < 1) return -21;
  if (L > 64 || V > 64) return -22;
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* xlx_version(void) { return "xlxmert_b200 0.1.0 (sm_100a)"; }

const char* xlx_strerror(int32_t code) {
  switch (code) {
    case 0: return "ok";
    case -1: return "invalid GEMM arguments";
    case -2: return "misaligned pointer or leading dimension (16-byte granularity required)";
    case -3: return "GEMM tile does not fit shared memory";
    case -4: return "unsupported hidden size for LayerNorm (128/256/512/768/1024)";
    case -5: return "attention sequence length out of range (1..64)";
    case -6: return "unsupported convolution geometry for the implicit-GEMM path";
    case -10: return "cuTensorMapEncodeTiled unavailable (driver too old?)";
    case -11: return "cuTensorMapEncodeTiled failed";
    case -20: return "unsupported xlx_dims";
    case -21: return "batch / sequence sizes must be positive";
    case -22: return "sequence length > 64 is not supported by the attention kernel";
    case -23: return "workspace too small";
    case -24: return "null pointer argument";
    case -25: return "invalid backward stage mask";
    default: return code > 0 ? "CUDA error (see cudaGetErrorString)" : "unknown error";
  }
}

int32_t xlx_dropout_mask(uint64_t seed, uint32_t site, float p, int32_t kind, int64_t rows, int32_t cols, float* out,
                         void* stream) {
  if (!out) return -24;
  if (rows < 1 || cols < 1 || !(p >= 0.f && p < 1.f)) return -21;
  XLX_TRY(ensure_device(out));
  DropSite d = make_site(seed, site, p);
  if (!d.threshold) { d.scale = 1.f; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return kind == 0 ? dropout_mask_hidden(d, static_cast<size_t>(rows), cols, out, st)
                   : dropout_mask_probs(d, static_cast<size_t>(rows), cols, out, st);
}
int32_t xlx_dropout_site(const xlx_dims* d, int32_t what, int32_t block, int32_t direction) {
  if (!dims_ok(d)) return -20;
  const int n_att = d->l_layers + d->r_layers + 3 * d->x_layers, n_ffn = d->l_layers + d->r_layers + 2 * d->x_layers;
  if (what == 0 || what == 1) {
    if (block < 0 || block >= n_att || direction < 0 || direction > 1) return -1;
    return static_cast<int32_t>(what == 0 ? site_probs(block, direction) : site_att_out(block));
  }
  if (what == 2) {
    if (block < 0 || block >= n_ffn) return -1;
    return static_cast<int32_t>(site_ffn_out(n_att, block));
  }
  return -1;
}

int64_t xlx_launch_count(void) { return gemm_launch_count() + aux_launch_count() + xlx_generator_launch_count() + xlx_optim_launch_count(); }
int64_t xlx_gemm_launch_count(void) { return gemm_launch_count(); }
void xlx_profile_gemm_begin(void) { gemm_timing_begin(); }
int32_t xlx_profile_gemm_end(double* total_ms, double* total_flops, int64_t* launches) {
  long long n = 0;
  int rc = gemm_timing_end(total_ms, total_flops, &n);
  *launches = n;
  return rc;
}

int64_t xlx_encoder_num_params(const xlx_dims* d) {
  return dims_ok(d) ? static_cast<int64_t>(slot_table(d).elems.size()) : -20;
}
int64_t xlx_encoder_param_elems(const xlx_dims* d, int64_t slot) {
  if (!dims_ok(d)) return -20;
  SlotTable t = slot_table(d);
  return (slot < 0 || slot >= static_cast<int64_t>(t.elems.size())) ? -1 : t.elems[slot];
}
int64_t xlx_encoder_grad_offset(const xlx_dims* d, int64_t slot) {
  if (!dims_ok(d)) return -20;
  SlotTable t = slot_table(d);
  return (slot < 0 || slot >= static_cast<int64_t>(t.offset.size())) ? -1 : t.offset[slot];
}
int64_t xlx_encoder_grad_elems(const xlx_dims* d) { return dims_ok(d) ? slot_table(d).total : -20; }

size_t xlx_encoder_prep_bytes(const xlx_dims* d) { return dims_ok(d) ? prep_layout(d, nullptr).bytes : 0; }

int32_t xlx_encoder_prepare(const xlx_dims* d, const float* const* params, void* prep, void* stream) {
  if (!dims_ok(d)) return -20;
  if (!params || !prep) return -24;
  XLX_TRY(ensure_device(prep));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Prep p = prep_layout(d, prep);
  const int H = d->hidden, I = d->intermediate, F = d->feat_dim;
  std::vector<SplitJob> jobs;
  jobs.reserve(p.max_jobs);
  auto add = [&](const float* src, int R, int C, Split dst, int row0, Split dst_t, int ld_t, int col0_t) {
    SplitJob j;
    j.src = src; j.hi = dst.hi; j.lo = dst.lo; j.t_hi = dst_t.hi; j.t_lo = dst_t.lo;
    j.R = R; j.C = C; j.row0 = row0; j.ld_t = ld_t; j.col0_t = col0_t; j.tile0 = 0;
    jobs.push_back(j);
  };
  add(params[0], H, F, p.visn_w, 0, p.visn_w_t, H, 0);
  for (size_t blk = 0; blk < p.att.size(); ++blk) {
    const int s0 = att_slot(d, static_cast<int>(blk));
    const AttW& a = p.att[blk];
    // fused [3H, H] weight and its transpose [H, 3H]: q | k | v stacked along rows / columns
    for (int j = 0; j < 3; ++j) add(params[s0 + j], H, H, a.wqkv, j * H, a.wqkv_t, 3 * H, j * H);
    XLX_TRY(concat3_f32(params[s0 + 3], params[s0 + 4], params[s0 + 5], a.bqkv, H, st));
    add(params[s0 + 6], H, H, a.wo, 0, a.wo_t, H, 0);
  }
  for (size_t blk = 0; blk < p.ffn.size(); ++blk) {
    const int s0 = ffn_slot(d, static_cast<int>(blk));
    add(params[s0], I, H, p.ffn[blk].w1, 0, p.ffn[blk].w1_t, I, 0);
    add(params[s0 + 2], H, I, p.ffn[blk].w2, 0, p.ffn[blk].w2_t, H, 0);
  }
  XLX_TRY(split_batch(jobs.data(), static_cast<int>(jobs.size()), st));
  return 0;
}

size_t xlx_encoder_workspace_bytes(const xlx_dims* d, int32_t B, int32_t L, int32_t V, int32_t training) {
  if (check_common(d, B, L, V)) return 0;
  return make_plan(d, B, L, V, training != 0, nullptr).bytes;
}

int32_t xlx_encoder_fwd(const xlx_dims* d, const float* const* params, const void* prep, int32_t B, int32_t L,
                        int32_t V, const float* lang_in, const float* lang_mask, const float* visual_feats,
                        const float* visual_pos, const float* vis_mask, float* lang_out, float* vis_out,
                        float* lang_hidden, float* vis_hidden, void* workspace, size_t workspace_bytes,
                        int32_t training, int32_t passes, const xlx_dropout* dropout, void* stream) {
  XLX_TRY(check_common(d, B, L, V));
  if (!params || !prep || !lang_in || !visual_feats || !visual_pos || !lang_out || !vis_out || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  XLX_TRY(ensure_device(workspace));
  Run r;
  r.d = d; r.params = params; r.passes = passes; r.st = static_cast<cudaStream_t>(stream);
  r.lmask = lang_mask; r.vmask = vis_mask;
  if (dropout) {
    if (!(dropout->p_hidden >= 0.f && dropout->p_hidden < 1.f && dropout->p_attn >= 0.f && dropout->p_attn < 1.f)) return -1;
    if ((dropout->p_hidden > 0.f || dropout->p_attn > 0.f) && !training) return -1;   // the backward must replay it
    r.drop.p_hidden = dropout->p_hidden; r.drop.p_attn = dropout->p_attn; r.drop.seed = dropout->seed;
  }
  r.n_att = d->l_layers + d->r_layers + 3 * d->x_layers;
  r.prep = prep_layout(d, const_cast<void*>(prep));
  r.plan = make_plan(d, B, L, V, training != 0, workspace);
  const Plan& p = r.plan;
  if (p.bytes > workspace_bytes) return -23;
  const int H = p.H, F = p.F, Ml = p.Ml, Mv = p.Mv;
  const int nl = d->l_layers, nr = d->r_layers, nx = d->x_layers;
  const size_t lh = static_cast<size_t>(Ml) * H, vh = static_cast<size_t>(Mv) * H;
  auto hidden_ptr = [&](float* base, int idx, size_t stride, bool last, float* final_out) -> float* {
    if (base) return base + idx * stride;
    return last ? final_out : nullptr;
  };

  // language-side blocks go to the side stream (rl), vision-side and joint blocks stay on the caller's (r)
  SideStream* side = side_stream();
  Run rl = r;
  if (side) rl.st = side->st;
  XLX_TRY(fork_to(side, r.st));
  // ---- visual feature encoder (HF:476-484): (LN(W·feat + b) + LN(Wp·pos + bp)) / 2
  XLX_TRY(split_f32(visual_feats, p.feats, static_cast<size_t>(Mv) * F, r.st));
  {
    GemmEpilogue e;
    e.bias = P(r, 1); e.out_f32 = p.y1; e.ld_out = H;
    XLX_TRY(linear(r, p.feats, Mv, F, r.prep.visn_w, H, e));
    XLX_TRY(box_linear_fwd(visual_pos, P(r, 4), P(r, 5), Mv, H, p.y2, r.st));
    XLX_TRY(layernorm_fwd(p.y2, P(r, 6), P(r, 7), d->ln_eps, Mv, H, 0.5f, nullptr, Split(), p.tbox,
                          p.training ? p.vstats + 2 * Mv : nullptr, p.training ? p.vstats + 3 * Mv : nullptr, r.st));
    XLX_TRY(layernorm_fwd(p.y1, P(r, 2), P(r, 3), d->ln_eps, Mv, H, 0.5f, p.tbox, p.vis0, nullptr,
                          p.training ? p.vstats : nullptr, p.training ? p.vstats + Mv : nullptr, r.st,
                          r.hidden_site(DROP_SITE_VISN)));      // dropout((x + y) / 2), HF:482
  }
  XLX_TRY(split_f32(lang_in, p.lang0, lh, rl.st));

  // ---- language layers (HF:524-529) alongside the vision layers (HF:532-537)
  for (int i = 0; i < nl; ++i) {
    XLX_TRY(att_self_fwd(rl, i, L, lang_mask, nullptr));
    XLX_TRY(ffn_fwd(rl, i, hidden_ptr(lang_hidden, i, lh, nx == 0 && i == nl - 1, lang_out)));
  }
  for (int i = 0; i < nr; ++i) {
    XLX_TRY(att_self_fwd(r, nl + i, V, vis_mask, nullptr));
    XLX_TRY(ffn_fwd(r, nl + i, hidden_ptr(vis_hidden, i, vh, nx == 0 && i == nr - 1, vis_out)));
  }
  // ---- cross-modality layers (HF:540-552): cross (joint) → self → FFN (per modality, side by side)
  for (int k = 0; k < nx; ++k) {
    const int a0 = nl + nr + 3 * k, f0 = nl + nr + 2 * k;
    const bool last = (k == nx - 1);
    XLX_TRY(join_from(side, r.st));
    XLX_TRY(att_cross_fwd(r, a0));
    XLX_TRY(fork_to(side, r.st));
    XLX_TRY(att_self_fwd(rl, a0 + 1, L, lang_mask, nullptr));
    XLX_TRY(ffn_fwd(rl, f0, hidden_ptr(lang_hidden, nl + k, lh, last, lang_out)));
    XLX_TRY(att_self_fwd(r, a0 + 2, V, vis_mask, nullptr));
    XLX_TRY(ffn_fwd(r, f0 + 1, hidden_ptr(vis_hidden, nr + k, vh, last, vis_out)));
  }
  XLX_TRY(join_from(side, r.st));
  // final outputs: when hidden states were requested the last state was written there — copy it out
  const int n_lang_states = nl + nx, n_vis_states = nr + nx;
  if (lang_hidden && n_lang_states > 0) {
    cudaError_t e = cudaMemcpyAsync(lang_out, lang_hidden + (n_lang_states - 1) * lh, lh * 4, cudaMemcpyDeviceToDevice, r.st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (vis_hidden && n_vis_states > 0) {
    cudaError_t e = cudaMemcpyAsync(vis_out, vis_hidden + (n_vis_states - 1) * vh, vh * 4, cudaMemcpyDeviceToDevice, r.st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (n_lang_states == 0) XLX_TRY(unsplit_f32(p.lang0, lang_out, lh, r.st));
  if (n_vis_states == 0) XLX_TRY(unsplit_f32(p.vis0, vis_out, vh, r.st));
  return 0;
}

int32_t xlx_encoder_bwd(const xlx_dims* d, const float* const* params, const void* prep, int32_t B, int32_t L,
                        int32_t V, const float* visual_pos, const float* d_lang_out, const float* d_vis_out,
                        float* d_lang_in, float* d_visual_feats, float* grads, void* workspace,
                        size_t workspace_bytes, int32_t passes, int32_t stages, const xlx_dropout* dropout,
                        void* language_done_event, void* stream) {
  XLX_TRY(check_common(d, B, L, V));
  if (!params || !prep || !visual_pos || !d_lang_in || !grads || !workspace) return -24;
  if (passes != 1 && passes != 3) return -1;
  if (stages <= 0 || stages > XLX_BWD_ALL) return -25;
  XLX_TRY(ensure_device(workspace));
  Run r;
  r.d = d; r.params = params; r.passes = passes; r.st = static_cast<cudaStream_t>(stream);
  r.lmask = nullptr; r.vmask = nullptr;
  if (dropout) {
    if (!(dropout->p_hidden >= 0.f && dropout->p_hidden < 1.f && dropout->p_attn >= 0.f && dropout->p_attn < 1.f)) return -1;
    r.drop.p_hidden = dropout->p_hidden; r.drop.p_attn = dropout->p_attn; r.drop.seed = dropout->seed;
  }
  r.n_att = d->l_layers + d->r_layers + 3 * d->x_layers;
  r.prep = prep_layout(d, const_cast<void*>(prep));
  r.plan = make_plan(d, B, L, V, true, workspace);
  const Plan& p = r.plan;
  if (p.bytes > workspace_bytes) return -23;
  XLX_TRY(gemm_splitk_ws_reset(p.sc[0].splitk, r.st));     // arrival counters of the folded split-K (once per call;
  XLX_TRY(gemm_splitk_ws_reset(p.sc[1].splitk, r.st));     //  the side stream forks from r.st after this point)
  Deferred def;
  static const bool defer_on = [] { const char* e = getenv("XLX_DEFER_FINISH"); return !(e && e[0] == '0'); }();
  def.arena = p.part_arena; def.cap = defer_on ? p.part_arena_floats : 0;
  Bwd bw;
  bw.r = &r; bw.sc = &p.sc[0]; bw.grads = grads; bw.slots = slot_table(d); bw.def = &def;
  // language-side blocks: side stream + their own temporaries (falls back to the caller's stream and set 0)
  SideStream* side = side_stream();
  Run rl = r;
  Bwd bwl = bw;
  if (side) { rl.st = side->st; bwl.r = &rl; bwl.sc = &p.sc[1]; bwl.lane = 1; }
  const int H = p.H, F = p.F, Ml = p.Ml, Mv = p.Mv;
  const int nl = d->l_layers, nr = d->r_layers, nx = d->x_layers;
  const size_t lh = static_cast<size_t>(Ml) * H, vh = static_cast<size_t>(Mv) * H;

  // gradient state: [language rows | vision rows], ping-pong between dA and dB (it lives in the workspace, so the
  // stages of one backward may be issued by separate calls; after the cross-modality stage the roles have swapped
  // x_layers times)
  float* cur = p.dA;
  float* nxt = p.dB;
  if (stages & XLX_BWD_CROSS) {
    auto load_grad = [&](float* dst, const float* src, size_t n) -> int {
      cudaError_t e = src ? cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, r.st)
                          : cudaMemsetAsync(dst, 0, n * 4, r.st);
      return e == cudaSuccess ? 0 : static_cast<int>(e);
    };
    XLX_TRY(load_grad(cur, d_lang_out, lh));
    XLX_TRY(load_grad(cur + lh, d_vis_out, vh));
    for (int k = nx - 1; k >= 0; --k) {
      const int a0 = nl + nr + 3 * k, f0 = nl + nr + 2 * k;
      // An output without upstream gradient (d_lang_out == NULL: the vis_mask task; d_vis_out == NULL: word_mask and
      // matched) leaves the LAST cross-modality layer's self-attention + FFN of that modality outside the graph: the
      // reference's autograd gives those parameters no gradient at all and AdamW skips them.  Their
      // blocks are not run, their arena slots are not written (the caller reports None), and the zero-filled rows of
      // `cur` are exactly the gradient the cross-attention block below has to see.
      const bool skip_lang = (k == nx - 1) && !d_lang_out, skip_vis = (k == nx - 1) && !d_vis_out;
      XLX_TRY(fork_to(side, r.st));
      if (!skip_lang) {
        XLX_TRY(ffn_bwd(bwl, f0, cur, nxt));                        // language chain
        XLX_TRY(att_self_bwd(bwl, a0 + 1, L, nxt, cur));
      }
      if (!skip_vis) {
        XLX_TRY(ffn_bwd(bw, f0 + 1, cur + lh, nxt + lh));           // vision chain
        XLX_TRY(att_self_bwd(bw, a0 + 2, V, nxt + lh, cur + lh));
      }
      XLX_TRY(join_from(side, r.st));
      XLX_TRY(att_cross_bwd(bw, a0, cur, nxt));                     // joint
      float* t = cur; cur = nxt; nxt = t;
    }
  } else if (nx & 1) {
    cur = p.dB; nxt = p.dA;
  }
  // the language stack runs beside the vision stack / visual feature encoder when the same call asks for both
  const bool lang_beside = side && (stages & XLX_BWD_LANGUAGE) && (stages & (XLX_BWD_VISION | XLX_BWD_VISN_FC));
  if (stages & XLX_BWD_LANGUAGE) {
    const Bwd& b = lang_beside ? bwl : bw;
    if (lang_beside) XLX_TRY(fork_to(side, r.st));
    for (int i = nl - 1; i >= 0; --i) {
      XLX_TRY(ffn_bwd(b, i, cur, nxt));
      XLX_TRY(att_self_bwd(b, i, L, nxt, cur));
    }
    // language embedding gradient
    XLX_CUDA(cudaMemcpyAsync(d_lang_in, cur, lh * 4, cudaMemcpyDeviceToDevice, b.r->st));
    if (language_done_event) {
      // the language range of the arena is complete once this stream's deferred column sums are finished: do that
      // now, on this stream, and tell the caller — it may start reducing the range while the vision stack computes
      std::vector<FinishJob>& lj = def.jobs[b.lane];
      if (!lj.empty()) {
        XLX_TRY(colsum_finish_batched(lj.data(), static_cast<int>(lj.size()), b.r->st));
        lj.clear();
      }
      XLX_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(language_done_event), b.r->st));
    }
  }
  if (stages & XLX_BWD_VISION) {
    for (int i = nr - 1; i >= 0; --i) {
      XLX_TRY(ffn_bwd(bw, nl + i, cur + lh, nxt + lh));
      XLX_TRY(att_self_bwd(bw, nl + i, V, nxt + lh, cur + lh));
    }
  }
  if (stages & XLX_BWD_VISN_FC) {
    // visual feature encoder backward
    const float* dvis = cur + lh;
    const BwdScratch& c = p.sc[0];
    int nblk = 0;
    // box branch: d(0.5·LN_b(y2))
    const DropSite vdrop = r.hidden_site(DROP_SITE_VISN);   // the forward dropped (x + y) / 2: both branches see dvis ∘ mask
    XLX_TRY(layernorm_bwd(dvis, 0.5f, p.y2, P(r, 6), p.vstats + 2 * Mv, p.vstats + 3 * Mv, Mv, H, p.dy2, Split(),
                          c.part, &nblk, r.st, vdrop));
    float* o2[2] = {bw.G(6), bw.G(7)};
    XLX_TRY(colsum_finish(c.part, 2, nblk, H, o2, 0, r.st));
    XLX_TRY(box_linear_bwd(p.dy2, visual_pos, Mv, H, c.part, bw.G(4), bw.G(5), r.st));
    // feature branch: d(0.5·LN_v(y1))
    XLX_TRY(layernorm_bwd(dvis, 0.5f, p.y1, P(r, 2), p.vstats, p.vstats + Mv, Mv, H, nullptr, c.dy_s, c.part, &nblk, r.st,
                          vdrop));
    float* o1[3] = {bw.G(2), bw.G(3), bw.G(1)};      // LN affine grads + visn_fc bias grad (Σ_rows of the LN-input grad)
    XLX_TRY(colsum_finish(c.part, 3, nblk, H, o1, 0, r.st));
    XLX_TRY(wgrad(r, c.splitk, c.dy_s, Mv, H, p.feats, F, bw.G(0)));
    if (d_visual_feats) {
      GemmEpilogue e;
      e.out_f32 = d_visual_feats; e.ld_out = F;
      XLX_TRY(dgrad(r, c.dy_s, Mv, H, r.prep.visn_w_t, F, e));
    }
  }
  if (lang_beside) XLX_TRY(join_from(side, r.st));
  // every stream of this call has been joined: finish all deferred column sums in one (or two) launches
  for (int lane = 0; lane < 2; ++lane)
    if (!def.jobs[lane].empty())
      XLX_TRY(colsum_finish_batched(def.jobs[lane].data(), static_cast<int>(def.jobs[lane].size()), r.st));
  return 0;
}

// Byte offset, inside a TRAINING workspace, of the soft-max probabilities attention block `att_block` (plan order)
// saved for the backward: [B, heads, Sq, Sk] fp32, before dropout.  second = 1: the vision-query half of a cross block.
// This is what HF returns as language_attentions / vision_attentions / cross_encoder_attentions (HF:506-565) in eval mode.
int64_t xlx_encoder_probs_offset(const xlx_dims* d, int32_t B, int32_t L, int32_t V, int32_t att_block, int32_t second) {
  if (check_common(d, B, L, V)) return -1;
  Plan p = make_plan(d, B, L, V, true, nullptr);
  if (att_block < 0 || att_block >= static_cast<int>(p.att.size())) return -1;
  const float* ptr = second ? p.att[att_block].probs2 : p.att[att_block].probs;
  if (!ptr) return -1;
  return static_cast<int64_t>(reinterpret_cast<const char*>(ptr) - static_cast<const char*>(nullptr));
}

// Element range [*offset, *offset + *elems) of the gradient arena completed by one backward stage.
int32_t xlx_encoder_grad_stage_range(const xlx_dims* d, int32_t stage, int64_t* offset, int64_t* elems) {
  if (!dims_ok(d)) return -20;
  if (!offset || !elems) return -24;
  SlotTable t = slot_table(d);
  const int nlr_slots = (N_ATT + N_FFN);
  const int s_l0 = N_VISN, s_r0 = N_VISN + d->l_layers * nlr_slots, s_x0 = s_r0 + d->r_layers * nlr_slots;
  const int s_end = static_cast<int>(t.elems.size());
  int a, b;
  switch (stage) {
    case XLX_BWD_CROSS: a = s_x0; b = s_end; break;
    case XLX_BWD_VISION: a = s_r0; b = s_x0; break;
    case XLX_BWD_LANGUAGE: a = s_l0; b = s_r0; break;
    case XLX_BWD_VISN_FC: a = 0; b = s_l0; break;
    default: return -25;
  }
  *offset = a < s_end ? t.offset[a] : t.total;
  *elems = (b < s_end ? t.offset[b] : t.total) - *offset;
  return 0;
}

}  // extern "C"
