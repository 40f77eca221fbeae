// Counter-based dropout for the fused kernels (training mode of the reference: nn.Dropout(0.1) at HF
// modeling_lxmert.py:189,236,282,344,474; the reference trains under model.train(), lxmert_pretrain.py:271).
//
// A mask is a pure function of (seed, site, element index): Philox4x32-10 keyed by the 64-bit seed, counter =
// (group index lo, hi, site, 0).  The forward applies it inside the producing kernel, the backward regenerates it from the
// same triple — nothing is stored.  PyTorch's own Philox stream cannot be bit-matched from fused kernels,
// so equivalence with the reference is statistical; parity of the arithmetic AROUND the masks is tested exactly by
// materialising these masks (xlx_dropout_mask_*) and feeding them to the oracle.
//
// Sites (one per nn.Dropout call of the reference's forward):
//   0                      embeddings output                                   (HF:189,212)
//   1                      visual feature encoder output                       (HF:474,482)
//   16 + 4·a + 0 / 1       attention probabilities of attention block a (plan order); cross block: +0 language
//                          queries, +1 vision queries                          (HF:236,262-266)
//   16 + 4·a + 2           attention-output dense of block a                   (HF:282,286)
//   16 + 4·n_att + f       FFN-output dense of FFN block f                     (HF:344,348)
// Element → Philox group:
//   hidden sites:  element (row, col) of a [rows, H] matrix: group = (row·H + col) / 4, lane = col % 4   (H % 4 == 0)
//   prob sites:    element (r, j) of a [rows = B·heads·Sq, Sk] matrix: group = r·ceil(Sk/2) + j/2, lane = j % 2
#pragma once
#include <stdint.h>

namespace xlx {

struct DropoutCfg {
  float p_hidden = 0.f, p_attn = 0.f;
  uint64_t seed = 0;
  bool hidden_on() const { return p_hidden > 0.f; }
  bool attn_on() const { return p_attn > 0.f; }
};

// kernel-side description of one dropout application
struct DropSite {
  uint32_t seed_lo = 0, seed_hi = 0, site = 0;
  uint32_t threshold = 0;   // keep iff random u32 >= threshold (threshold = p·2^32); 0 = dropout off
  float scale = 1.f;        // 1 / (1 − p)
};

inline DropSite make_site(uint64_t seed, uint32_t site, float p) {
  DropSite d;
  if (p > 0.f) {
    d.seed_lo = static_cast<uint32_t>(seed);
    d.seed_hi = static_cast<uint32_t>(seed >> 32);
    d.site = site;
    double t = static_cast<double>(p) * 4294967296.0;
    d.threshold = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
    d.scale = 1.0f / (1.0f - p);
  }
  return d;
}

enum : uint32_t { DROP_SITE_EMB = 0, DROP_SITE_VISN = 1, DROP_SITE_BLOCK0 = 16 };
inline uint32_t site_probs(int att_block, int direction) { return DROP_SITE_BLOCK0 + 4u * att_block + direction; }
inline uint32_t site_att_out(int att_block) { return DROP_SITE_BLOCK0 + 4u * att_block + 2u; }
inline uint32_t site_ffn_out(int n_att, int ffn_block) { return DROP_SITE_BLOCK0 + 4u * n_att + ffn_block; }

#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// four random words of Philox group `group` at this site
__device__ __forceinline__ uint4 drop_words(const DropSite& d, uint64_t group) {
  return philox4x32_10(static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32), d.site, 0u, d.seed_lo, d.seed_hi);
}
// multiplier (0 or 1/(1−p)) for one random word
__device__ __forceinline__ float drop_mul(const DropSite& d, uint32_t word) { return word >= d.threshold ? d.scale : 0.f; }
// multipliers of the four consecutive columns col … col+3 (col % 4 == 0) of row `row` of a [rows, H] hidden matrix
__device__ __forceinline__ float4 drop_hidden4(const DropSite& d, size_t row, int H, int col) {
  const uint4 w = drop_words(d, (static_cast<uint64_t>(row) * H + col) >> 2);
  return make_float4(drop_mul(d, w.x), drop_mul(d, w.y), drop_mul(d, w.z), drop_mul(d, w.w));
}
// multipliers of the probability pair (r, j), (r, j+1) with j even of a [rows, Sk] matrix
__device__ __forceinline__ float2 drop_prob2(const DropSite& d, size_t r, int Sk, int j) {
  const uint4 w = drop_words(d, static_cast<uint64_t>(r) * ((Sk + 1) >> 1) + (j >> 1));
  return make_float2(drop_mul(d, w.x), drop_mul(d, w.y));
}
// multiplier of the single probability (r, j) — same value as the matching lane of drop_prob2
__device__ __forceinline__ float drop_prob1(const DropSite& d, size_t r, int Sk, int j) {
  const uint4 w = drop_words(d, static_cast<uint64_t>(r) * ((Sk + 1) >> 1) + (j >> 1));
  return drop_mul(d, (j & 1) ? w.y : w.x);
}
#endif

}  // namespace xlx
