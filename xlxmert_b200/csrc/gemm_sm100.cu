// tcgen05 / TMA / TMEM GEMM for sm_100a — see gemm_sm100.cuh for the contract.
#include "gemm_sm100.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "xlx_ptx.cuh"

namespace xlx {

namespace {

constexpr int BM = 128;            // UMMA M (one TMEM lane per output row)
constexpr int GEMM_THREADS = 192;  // 6 warps: TMA, MMA, 4 × epilogue
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct KParams {
  int M, N, K;
  int BN;
  int nparts;  // 1 (hi only) or 2 (hi + lo)
  int a_mn, b_mn;
  int num_stages;
  int tiles_m, tiles_n;
  uint32_t stage_bytes, a_part_bytes, b_part_bytes;
  uint32_t tmem_cols;
  GemmEpilogue epi;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  // d/dx [x·Φ(x)] = Φ(x) + x·φ(x)
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// One thread's slice of the epilogue: 32 consecutive columns of one output row.
__device__ __forceinline__ void epilogue_row(const KParams& P, int row, int n, int nvalid, uint32_t (&r)[32]) {
  const GemmEpilogue& E = P.epi;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * E.alpha;
  if (E.bias) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g * 4 < nvalid) {
        float4 b = __ldg(reinterpret_cast<const float4*>(E.bias + n + g * 4));
        v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
      }
    }
  }
  if (E.out_u) {
    float* dst = E.out_u + static_cast<size_t>(row) * E.ld_u + n;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (g * 4 < nvalid)
        *reinterpret_cast<float4*>(dst + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
  }
  if (E.flags & EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
  if (E.flags & EPI_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  if (E.flags & EPI_GELU_GRAD) {
    const float* src = E.u_in + static_cast<size_t>(row) * E.ld_u + n;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g * 4 < nvalid) {
        float4 u = __ldg(reinterpret_cast<const float4*>(src + g * 4));
        v[g * 4 + 0] *= gelu_erf_grad(u.x); v[g * 4 + 1] *= gelu_erf_grad(u.y);
        v[g * 4 + 2] *= gelu_erf_grad(u.z); v[g * 4 + 3] *= gelu_erf_grad(u.w);
      }
    }
  }
  if (E.addend) {
    const float* src = E.addend + static_cast<size_t>(row) * E.ld_addend + n;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g * 4 < nvalid) {
        float4 a = __ldg(reinterpret_cast<const float4*>(src + g * 4));
        v[g * 4 + 0] += a.x; v[g * 4 + 1] += a.y; v[g * 4 + 2] += a.z; v[g * 4 + 3] += a.w;
      }
    }
  }
  if (E.addend_hi) {
    const __nv_bfloat16* sh = E.addend_hi + static_cast<size_t>(row) * E.ld_addend + n;
    const __nv_bfloat16* sl = E.addend_lo + static_cast<size_t>(row) * E.ld_addend + n;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g * 8 < nvalid) {
        uint4 h = __ldg(reinterpret_cast<const uint4*>(sh + g * 8));
        uint4 l = __ldg(reinterpret_cast<const uint4*>(sl + g * 8));
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          v[g * 8 + 2 * t] += __uint_as_float(hh[t] << 16) + __uint_as_float(ll[t] << 16);
          v[g * 8 + 2 * t + 1] += __uint_as_float(hh[t] & 0xffff0000u) + __uint_as_float(ll[t] & 0xffff0000u);
        }
      }
    }
  }
  if (E.out_f32) {
    float* dst = E.out_f32 + static_cast<size_t>(row) * E.ld_out + n;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g * 4 < nvalid) {
        float4 o = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        if (E.flags & EPI_ACCUM) {
          float4 old = *reinterpret_cast<const float4*>(dst + g * 4);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + g * 4) = o;
      }
    }
  }
  if (E.out_hi) {
    __nv_bfloat16* dh = E.out_hi + static_cast<size_t>(row) * E.ld_split + n;
    __nv_bfloat16* dl = E.out_lo ? E.out_lo + static_cast<size_t>(row) * E.ld_split + n : nullptr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g * 8 < nvalid) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[g * 8 + 2 * t], h0, l0);
          split_bf16(v[g * 8 + 2 * t + 1], h1, l1);
          hw[t] = pack_bf16x2(h0, h1);
          lw[t] = pack_bf16x2(l0, l1);
        }
        *reinterpret_cast<uint4*>(dh + g * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (dl) *reinterpret_cast<uint4*>(dl + g * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
}

template <int BK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
            const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
            const KParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms need 1024B alignment

  const int num_tiles = P.tiles_m * P.tiles_n;
  const int nkb = (P.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapAhi);
    tma_prefetch_desc(&mapBhi);
    if (P.nparts == 2) {
      tma_prefetch_desc(&mapAlo);
      tma_prefetch_desc(&mapBlo);
    }
    for (int s = 0; s < P.num_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.tiles_n) * BM;
        const int n0 = (tile % P.tiles_n) * P.BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          mbar_arrive_expect_tx(full, P.stage_bytes);
          const uint32_t sA = smem_base + s * P.stage_bytes;
          const uint32_t sB = sA + P.nparts * P.a_part_bytes;
          const int k0 = kb * BK;
          for (int part = 0; part < P.nparts; ++part) {
            const CUtensorMap* ma = part ? &mapAlo : &mapAhi;
            const CUtensorMap* mb = part ? &mapBlo : &mapBhi;
            const uint32_t dA = sA + part * P.a_part_bytes;
            const uint32_t dB = sB + part * P.b_part_bytes;
            if (!P.a_mn) {
              tma_load_2d(dA, ma, full, k0, m0);
            } else {
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(dA + j * (BK * 128), ma, full, m0 + 64 * j, k0);
            }
            if (!P.b_mn) {
              tma_load_2d(dB, mb, full, k0, n0);
            } else {
              for (int j = 0; j < P.BN / 64; ++j) tma_load_2d(dB + j * (BK * 128), mb, full, n0 + 64 * j, k0);
            }
          }
          if (++s == P.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, P.BN, P.a_mn, P.b_mn);
      // K-major: rows of BK·2 bytes (64B or 128B swizzle), 8-row groups contiguous.
      constexpr uint32_t kSwzK = (BK == 64) ? UMMA_SWZ_128B : UMMA_SWZ_64B;
      constexpr uint32_t kSboK = 8 * BK * 2;
      constexpr uint32_t kStepK = 32;          // 16 bf16 along K inside the swizzled row
      // MN-major: 64-element (128B) MN atoms, one TMA box of BK rows each; 8-row K groups 1024B apart.
      constexpr uint32_t kLboMN = BK * 128;
      constexpr uint32_t kSboMN = 1024;
      constexpr uint32_t kStepMN = 16 * 128;   // 16 K rows
      int s = 0;
      uint32_t ph = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(smem_u32(&bar_tmem_empty[buf]), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * P.BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&bar_full[s]), ph);
          tc_fence_after();
          const uint32_t sA = smem_base + s * P.stage_bytes;
          const uint32_t sB = sA + P.nparts * P.a_part_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t offA = P.a_mn ? kk * kStepMN : kk * kStepK;
            const uint32_t offB = P.b_mn ? kk * kStepMN : kk * kStepK;
            const uint64_t dAhi = P.a_mn ? umma_smem_desc(sA + offA, kLboMN, kSboMN, UMMA_SWZ_128B)
                                         : umma_smem_desc(sA + offA, 16, kSboK, kSwzK);
            const uint64_t dBhi = P.b_mn ? umma_smem_desc(sB + offB, kLboMN, kSboMN, UMMA_SWZ_128B)
                                         : umma_smem_desc(sB + offB, 16, kSboK, kSwzK);
            const uint32_t first = (kb | kk) ? 1u : 0u;
            if (P.nparts == 2) {
              const uint64_t dAlo = P.a_mn ? umma_smem_desc(sA + P.a_part_bytes + offA, kLboMN, kSboMN, UMMA_SWZ_128B)
                                           : umma_smem_desc(sA + P.a_part_bytes + offA, 16, kSboK, kSwzK);
              const uint64_t dBlo = P.b_mn ? umma_smem_desc(sB + P.b_part_bytes + offB, kLboMN, kSboMN, UMMA_SWZ_128B)
                                           : umma_smem_desc(sB + P.b_part_bytes + offB, 16, kSboK, kSwzK);
              // small cross terms first, leading term last
              umma_bf16(tmem_d, dAlo, dBhi, idesc, first);
              umma_bf16(tmem_d, dAhi, dBlo, idesc, 1u);
              umma_bf16(tmem_d, dAhi, dBhi, idesc, 1u);
            } else {
              umma_bf16(tmem_d, dAhi, dBhi, idesc, first);
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));              // smem slot reusable once these MMAs retire
          if (kb == nkb - 1) umma_commit(smem_u32(&bar_tmem_full[buf]));
          if (++s == P.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      const int m0 = (tile / P.tiles_n) * BM;
      const int n0 = (tile % P.tiles_n) * P.BN;
      mbar_wait(smem_u32(&bar_tmem_full[buf]), acc_ph);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * P.BN;
      const int nchunks = P.BN / 32;
      for (int c = 0; c < nchunks; ++c) {
        const int n = n0 + c * 32;
        if (n >= P.N) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const bool last = (c == nchunks - 1) || (n + 32 >= P.N);
        if (last) {  // accumulator fully read: hand the TMEM buffer back before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
        }
        if (row < P.M) epilogue_row(P, row, n, min(32, P.N - n), r);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box0, box1, swz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box0 == o.box0 &&
           box1 == o.box1 && swz == o.swz;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box0); mix(k.box1); mix(k.swz);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;
std::atomic<long long> g_launches{0};

// optional per-launch timing (bench.py roofline): events bracket every GEMM on its own stream
struct TimedLaunch { cudaEvent_t e0, e1; double flops; int M, N, K, passes, a_mn, b_mn, flags, grid; };
bool g_timing = false;
std::vector<TimedLaunch> g_timed;

// bf16 row-major [outer, inner] with leading dimension ld (elements); box = box0 (inner) × box1 (outer).
int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box0,
             uint32_t box1, CUtensorMapSwizzle swz) {
  MapKey key{ptr, inner, outer, ld, box0, box1, static_cast<uint32_t>(swz)};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -11;
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(key, *out);
  return 0;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int BK>
int launch_bk(const GemmProblem& p, cudaStream_t stream) {
  static int num_sms = 0;
  static bool attr_set = false;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  KParams P;
  memset(&P, 0, sizeof(P));
  P.M = p.M; P.N = p.N; P.K = p.K;
  int BN = 256;
  if (p.N <= 64) BN = 64; else if (p.N <= 128) BN = 128;
  int force_bn = env_int("XLX_GEMM_BN", 0);
  if (force_bn) BN = force_bn;
  P.BN = BN;
  P.nparts = (p.passes == 3) ? 2 : 1;
  P.a_mn = p.a.mn_major; P.b_mn = p.b.mn_major;
  P.a_part_bytes = BM * BK * 2;
  P.b_part_bytes = BN * BK * 2;
  P.stage_bytes = P.nparts * (P.a_part_bytes + P.b_part_bytes);
  int stages = (SMEM_LIMIT - 2048 - 1024) / static_cast<int>(P.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return -3;
  P.num_stages = stages;
  P.tiles_m = (p.M + BM - 1) / BM;
  P.tiles_n = (p.N + BN - 1) / BN;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * BN)) cols <<= 1;
  P.tmem_cols = cols;
  P.epi = p.epi;

  CUtensorMap mAhi, mAlo, mBhi, mBlo;
  const CUtensorMapSwizzle swzK = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  auto mk = [&](CUtensorMap* m, const GemmOperand& o, const __nv_bfloat16* ptr, int rows, int box_rows) -> int {
    const int kext = o.kext ? o.kext : p.K;
    if (!o.mn_major) return make_map(m, ptr, kext, rows, o.ld, BK, box_rows, swzK);
    return make_map(m, ptr, rows, kext, o.ld, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  const int a_rows = p.a.rows ? p.a.rows : p.M, b_rows = p.b.rows ? p.b.rows : p.N;
  if ((rc = mk(&mAhi, p.a, p.a.hi, a_rows, BM))) return rc;
  if ((rc = mk(&mBhi, p.b, p.b.hi, b_rows, BN))) return rc;
  if (P.nparts == 2) {
    if ((rc = mk(&mAlo, p.a, p.a.lo, a_rows, BM))) return rc;
    if ((rc = mk(&mBlo, p.b, p.b.lo, b_rows, BN))) return rc;
  } else {
    mAlo = mAhi; mBlo = mBhi;
  }
  const int num_tiles = P.tiles_m * P.tiles_n;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  const size_t smem = static_cast<size_t>(stages) * P.stage_bytes + 1024;
  TimedLaunch tl{};
  if (g_timing) {
    cudaEventCreate(&tl.e0);
    cudaEventCreate(&tl.e1);
    tl.flops = 2.0 * p.M * p.N * p.K;
    tl.M = p.M; tl.N = p.N; tl.K = p.K; tl.passes = p.passes; tl.a_mn = p.a.mn_major; tl.b_mn = p.b.mn_major;
    tl.flags = p.epi.flags | (p.epi.out_f32 ? 256 : 0) | (p.epi.out_hi ? 512 : 0) | (p.epi.addend_hi ? 1024 : 0) |
               (p.epi.out_u ? 2048 : 0) | (p.epi.addend ? 4096 : 0);
    tl.grid = grid;
    cudaEventRecord(tl.e0, stream);
  }
  gemm_kernel<BK><<<grid, GEMM_THREADS, smem, stream>>>(mAhi, mAlo, mBhi, mBlo, P);
  if (g_timing) {
    cudaEventRecord(tl.e1, stream);
    g_timed.push_back(tl);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace

long long gemm_launch_count() { return g_launches.load(); }

void gemm_timing_begin() {
  g_timed.clear();
  g_timing = true;
}
int gemm_timing_end(double* total_ms, double* total_flops, long long* launches) {
  g_timing = false;
  double ms = 0, fl = 0;
  FILE* log = nullptr;
  if (const char* path = getenv("XLX_GEMM_LOG")) log = fopen(path, "w");
  if (log) fprintf(log, "M,N,K,passes,a_mn,b_mn,epi,grid,us\n");
  for (auto& t : g_timed) {
    cudaError_t e = cudaEventSynchronize(t.e1);
    if (e != cudaSuccess) return static_cast<int>(e);
    float m = 0;
    cudaEventElapsedTime(&m, t.e0, t.e1);
    ms += m; fl += t.flops;
    if (log) fprintf(log, "%d,%d,%d,%d,%d,%d,%d,%d,%.2f\n", t.M, t.N, t.K, t.passes, t.a_mn, t.b_mn, t.flags, t.grid, m * 1e3);
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  if (log) fclose(log);
  *total_ms = ms; *total_flops = fl; *launches = static_cast<long long>(g_timed.size());
  g_timed.clear();
  return 0;
}

int gemm_launch(const GemmProblem& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return -1;
  if (p.passes != 1 && p.passes != 3) return -1;
  if (!p.a.hi || !p.b.hi) return -1;
  if (p.passes == 3 && (!p.a.lo || !p.b.lo)) return -1;
  if ((p.a.ld % 8) || (p.b.ld % 8) || (p.N % 8)) return -2;  // TMA: 16-byte global strides; epilogue: 16B vectors
  if ((reinterpret_cast<uintptr_t>(p.a.hi) | reinterpret_cast<uintptr_t>(p.b.hi) |
       reinterpret_cast<uintptr_t>(p.a.lo) | reinterpret_cast<uintptr_t>(p.b.lo)) & 15)
    return -2;
  const GemmEpilogue& E = p.epi;
  if ((E.out_f32 && (E.ld_out % 4)) || (E.out_hi && (E.ld_split % 8)) ||
      ((E.addend || E.addend_hi) && (E.ld_addend % 8)) || ((E.out_u || E.u_in) && (E.ld_u % 4)))
    return -2;
  if ((E.flags & EPI_GELU_GRAD) && !E.u_in) return -1;
  if (E.addend_hi && !E.addend_lo) return -1;
  static const int bk = env_int("XLX_GEMM_BK", 32);
  return bk == 64 ? launch_bk<64>(p, stream) : launch_bk<32>(p, stream);
}

}  // namespace xlx
