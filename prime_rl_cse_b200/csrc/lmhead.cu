// Fused LM head (forward statistics, saved-logits and recompute backward) and the plain bring-up GEMM, host side:
// TMA tensor maps, work decomposition, kernel launches, C ABI.  The device code is gemm_engine.cuh.
#include <stdlib.h>

#include "gemm_engine.cuh"

namespace prl {

// ------------------------------------------------------------------------------------------------
// Tuning knobs.  Every knob has a compiled-in default (the measured winner); the environment variable
// of the same name is read ONCE, the first time the library needs a knob; prl_set_knob() changes a value afterwards
// (tests and the tuning scripts under tools/).  PRL_CTA_GROUP, PRL_MSUB and PRL_MAX_WORKERS select kernel variants whose
// launch attributes are cached on first use: set them before the first call.  Knobs marked `debug` exist only in builds with
// -DPRL_DEBUG_KNOBS (make debug): the profiling aids behind them (which produce WRONG results) are not compiled into the
// release kernels.
// ------------------------------------------------------------------------------------------------
enum KnobId {
  K_CTA_GROUP, K_MSUB, K_EPI_TMA, K_L2_PROMO, K_K_SPLITS, K_GROUP_M, K_TILE_STAGES, K_SWEEP_GROUP_M, K_FWD_STAGES, K_DZ_STAGES, K_VCHUNK,
  K_HINT0,  // 18 hints: family (FWD, FWDS, DZ, DH, DW, GEMM) x operand (A, B, C); -1 = the family's compiled-in default
  K_HINT_END = K_HINT0 + 18,
#ifdef PRL_DEBUG_KNOBS
  K_SPLIT_GROUP = K_HINT_END, K_MAX_WORKERS, K_WAIT_BACKOFF, K_DEBUG_NOSTORE, K_DEBUG_SKIPB, K_DEBUG_MN_LBO, K_DEBUG_MN_SBO, K_DEBUG_MN_KSTEP,
  K_COUNT
#else
  K_COUNT = K_HINT_END
#endif
};
static const char* const kHintFamilies[6] = {"FWD", "FWDS", "DZ", "DH", "DW", "GEMM"};

struct KnobTable {
  char names[K_COUNT][32];
  long values[K_COUNT];
  KnobTable() {
    auto def = [&](int id, const char* name, long dflt) {
      snprintf(names[id], sizeof(names[id]), "%s", name);
      const char* e = getenv(name);
      values[id] = e ? strtol(e, nullptr, 0) : dflt;
    };
    def(K_CTA_GROUP, "PRL_CTA_GROUP", 2);
    def(K_MSUB, "PRL_MSUB", 0);
    def(K_EPI_TMA, "PRL_EPI_TMA", 1);
    def(K_L2_PROMO, "PRL_L2_PROMO", 256);
    def(K_K_SPLITS, "PRL_K_SPLITS", 0);
    def(K_GROUP_M, "PRL_GROUP_M", 0);
    def(K_TILE_STAGES, "PRL_TILE_STAGES", 0);
    def(K_SWEEP_GROUP_M, "PRL_SWEEP_GROUP_M", 32);
    def(K_FWD_STAGES, "PRL_FWD_STAGES", 5);
    def(K_DZ_STAGES, "PRL_DZ_STAGES", 5);
    def(K_VCHUNK, "PRL_VCHUNK", 16384);
    for (int f = 0; f < 6; ++f)
      for (int o = 0; o < 3; ++o) {
        char name[32];
        snprintf(name, sizeof(name), "PRL_HINT_%s_%c", kHintFamilies[f], "ABC"[o]);
        def(K_HINT0 + 3 * f + o, name, -1);
      }
#ifdef PRL_DEBUG_KNOBS
    def(K_SPLIT_GROUP, "PRL_SPLIT_GROUP", 0);
    def(K_MAX_WORKERS, "PRL_MAX_WORKERS", 0);
    def(K_WAIT_BACKOFF, "PRL_WAIT_BACKOFF", 0);
    def(K_DEBUG_NOSTORE, "PRL_DEBUG_NOSTORE", 0);
    def(K_DEBUG_SKIPB, "PRL_DEBUG_SKIPB", 0);
    def(K_DEBUG_MN_LBO, "PRL_DEBUG_MN_LBO", BLOCK_K * 128);
    def(K_DEBUG_MN_SBO, "PRL_DEBUG_MN_SBO", 1024);
    def(K_DEBUG_MN_KSTEP, "PRL_DEBUG_MN_KSTEP", UMMA_K * 128);
#endif
  }
};
static KnobTable& knobs() {
  static KnobTable t;  // C++11: initialised once, thread-safe
  return t;
}
static long knob(int id) { return knobs().values[id]; }

// CTAs per cluster: 1 = one CTA per 128x256 tile; 2 = a CTA pair (cta_group::2) per 256x256 tile; 4 = two pairs per
// 512x256 tile sharing B by TMA multicast
static int cta_group() {
  static int cg = 0;
  if (cg == 0) {
    const long v = knob(K_CTA_GROUP);
    cg = (v == 1) ? 1 : (v == 4 ? 4 : 2);
  }
  return cg;
}

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// L2 promotion of TMA requests (PRL_L2_PROMO = 0 none | 64 | 128 | 256; default 256)
static CUtensorMapL2promotion l2_promotion() {
  const long v = knob(K_L2_PROMO);
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
}

// matrix with `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements; 128B-swizzled boxes.
static int make_tmap_t(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
                       bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  const int64_t esz = f32 ? 4 : 2;
  if (!fn) return fail(PRL_ERR_CUDA, "%s", "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * esz) % 16 != 0)
    return fail(PRL_ERR_UNSUPPORTED, "%s", "TMA tensor needs a 16-byte aligned base and a row pitch that is a multiple of 16 bytes");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * esz)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PRL_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %s%lld)", "", (long long)r);
  return PRL_OK;
}

static int make_tmap(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  return make_tmap_t(map, base, inner, outer, ld, box_inner, box_outer, false);
}
static bool epi_tma_enabled() { return knob(K_EPI_TMA) != 0; }
// output map of the TMA-store epilogues: [rows, cols] row-major with pitch ld; fp32 boxes 32 x 32, bf16 boxes 64 x 32 (128 B rows)
static int make_output_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, bool f32) {
  return make_tmap_t(map, base, cols, rows, ld, f32 ? 32 : 64, 32, f32);
}
// operand description: K-major => stored [rows, K]; MN-major => stored [K, rows]
static int make_operand_map(CUtensorMap* map, const void* base, bool mn_major, int64_t rows, int64_t K, int64_t ld, int block_rows) {
  // B: a CTA of a pair stages half of the tile's rows; with two pairs per cluster it fetches only a quarter itself
  if (block_rows == BLOCK_N) block_rows = BLOCK_N / cta_group();
  if (!mn_major) return make_tmap(map, base, K, rows, ld, BLOCK_K, block_rows);
  return make_tmap(map, base, rows, K, ld, 64, BLOCK_K);
}

static void make_desc_defaults(GemmParams& p, bool a_mn, bool b_mn) {
  // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused (16 B); one UMMA_K step = +32 B.
  // MN-major SW128: 64-element MN slabs BLOCK_K*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
  //                 one UMMA_K step (16 k rows) = +2048 B.
#ifdef PRL_DEBUG_KNOBS
  const uint32_t mn_lbo = (uint32_t)knob(K_DEBUG_MN_LBO), mn_sbo = (uint32_t)knob(K_DEBUG_MN_SBO), mn_kstep = (uint32_t)knob(K_DEBUG_MN_KSTEP);
#else
  constexpr uint32_t mn_lbo = BLOCK_K * 128, mn_sbo = 1024, mn_kstep = UMMA_K * 128;
#endif
  p.a_desc_hi = a_mn ? ptx::make_smem_desc_hi_sw128(mn_lbo, mn_sbo) : ptx::make_smem_desc_hi_sw128(16, 1024);
  p.b_desc_hi = b_mn ? ptx::make_smem_desc_hi_sw128(mn_lbo, mn_sbo) : ptx::make_smem_desc_hi_sw128(16, 1024);
  p.a_kstep = a_mn ? mn_kstep : UMMA_K * 2;
  p.b_kstep = b_mn ? mn_kstep : UMMA_K * 2;
}

// M sub-tiles per CTA-pair tile: 1 = 256 x 256 tiles with double-buffered accumulators; 2 = 512 x 256 tiles (two sub-tiles
// share every B k-block: 25 % fewer operand bytes per MMA, no epilogue/mainloop overlap).  Only with PRL_CTA_GROUP=2.
// Measured: +12..16 % on the d_hidden / d_weight GEMMs (long K loops, cheap epilogue), -1 % on the
// forward / dlogits sweeps (K = H only, heavy epilogue) -> default 2 for the store-epilogue GEMMs, 1 for the sweeps.
// PRL_MSUB=1|2 forces one value everywhere.
static int msub(int epi) {
  static int forced = -1;
  if (forced < 0) forced = (int)knob(K_MSUB);
  if (cta_group() != 2) return 1;
  if (forced == 1 || forced == 2) return forced;
  return epi == EPI_STORE ? 2 : 1;
}

static void cluster_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int cg, int grid, int threads, int smem,
                                  cudaStream_t st) {
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// number of clusters that can be co-resident (persistent grid size): all SMs for CG <= 2, fewer for CG = 4 because
// clusters may not straddle GPCs
template <int CG>
static int resident_workers() {
  static int cached = 0;
  if (cached == 0) {
    int n = sm_count() / CG;
    if (CG > 2) {
      using C = Cfg<CG, 1, EPI_STORE>;
      auto kern = gemm_kernel<CG, 1, false, false, EPI_STORE>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[1];
      cluster_launch_config(cfg, attr, CG, sm_count() / CG * CG, C::kThreads, C::kSmem, nullptr);
      int q = 0;
      if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
      cudaGetLastError();
    }
#ifdef PRL_DEBUG_KNOBS
    const long cap = knob(K_MAX_WORKERS);  // diagnostic: use only the first `cap` clusters
    if (cap > 0 && cap < n) n = (int)cap;
#endif
    cached = n > 0 ? n : 1;
  }
  return cached;
}
static int workers() { return cta_group() == 4 ? resident_workers<4>() : (cta_group() == 2 ? resident_workers<2>() : resident_workers<1>()); }

// L2 eviction priorities per kernel family (a, b, c = the compiled-in defaults); each can be overridden for experiments with the
// knob PRL_HINT_<FAMILY>_<A|B|C> = 0 (no hint) | 1 (evict first) | 2 (evict last) | 3 (explicit evict-normal)
static uint64_t hint_value(long v) { return v == 1 ? ptx::kEvictFirst : (v == 2 ? ptx::kEvictLast : (v == 3 ? ptx::kEvictNormal : 0)); }
static void set_hints(GemmParams& p, const char* family, int a, int b, int c) {
  int f = 0;
  while (f < 5 && strcmp(kHintFamilies[f], family) != 0) ++f;
  const long ka = knob(K_HINT0 + 3 * f), kb = knob(K_HINT0 + 3 * f + 1), kc = knob(K_HINT0 + 3 * f + 2);
  p.hint_a = hint_value(ka >= 0 ? ka : a);
  p.hint_b = hint_value(kb >= 0 ? kb : b);
  p.hint_c = hint_value(kc >= 0 ? kc : c);
}

template <int CG, int MS, bool A_MN, bool B_MN, int EPI>
static int launch_gemm_cg(GemmParams p /* by value: the profiling knobs below edit the copy */, cudaStream_t st, const char* name) {
  using C = Cfg<CG, MS, EPI>;
  static bool attr_set[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  auto kern = gemm_kernel<CG, MS, A_MN, B_MN, EPI>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (p.num_items < 1) return PRL_OK;
#ifdef PRL_DEBUG_KNOBS
  if (knob(K_DEBUG_NOSTORE)) p.epi.flags |= ST_DEBUG_SKIP;
  if (knob(K_DEBUG_SKIPB)) p.epi.flags |= ST_DEBUG_SKIPB;
  if (knob(K_WAIT_BACKOFF)) p.epi.flags |= ST_WAIT_BACKOFF;
#endif
  int nw = resident_workers<CG>();
  if (nw > p.num_items) nw = p.num_items;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_launch_config(cfg, attr, CG, nw * CG, C::kThreads, C::kSmem, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return fail_cuda(e, name);
  return check_launch(name);
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const GemmParams& p, cudaStream_t st, const char* name) {
  if (cta_group() == 4) return launch_gemm_cg<4, 1, A_MN, B_MN, EPI>(p, st, name);
  if (cta_group() == 2 && msub(EPI) == 2) return launch_gemm_cg<2, 2, A_MN, B_MN, EPI>(p, st, name);
  if (cta_group() == 2) return launch_gemm_cg<2, 1, A_MN, B_MN, EPI>(p, st, name);
  return launch_gemm_cg<1, 1, A_MN, B_MN, EPI>(p, st, name);
}

static void set_tiling(GemmParams& p, int64_t M, int64_t N, int64_t K, int epi) {
  const int64_t bm = (int64_t)BLOCK_M * cta_group() * msub(epi);
  p.m_tiles = (int)((M + bm - 1) / bm);
  p.n_tiles = (int)((N + BLOCK_N - 1) / BLOCK_N);
  p.k_blocks = (int)((K + BLOCK_K - 1) / BLOCK_K);
  p.k_splits = 1;
  p.k_per_split = p.k_blocks;
}

// Two k-splits per output tile when that fills the machine markedly better (few output tiles, long K: d_hidden at small T x H).
// Only 1 or 2: with two partial sums the fp32 reduce-add into a zeroed accumulator gives the same bits in either arrival order.
static int pick_k_splits(int64_t M, int64_t N, int64_t K) {
  const long forced = knob(K_K_SPLITS);
  if (forced == 1 || forced == 2) return (int)forced;
  const int64_t bm = (int64_t)BLOCK_M * cta_group() * msub(EPI_STORE);
  const int64_t items = ((M + bm - 1) / bm) * ((N + BLOCK_N - 1) / BLOCK_N), w = workers();
  const int64_t kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (kb < 16) return 1;
  const double one = (double)((items + w - 1) / w), two = (double)((2 * items + w - 1) / w) / 2.0;
  return two < 0.97 * one ? 2 : 1;
}
static void set_k_splits(GemmParams& p, int ks) {
  if (ks <= 1) return;
  p.k_per_split = (p.k_blocks + ks - 1) / ks;
  p.k_splits = (p.k_blocks + p.k_per_split - 1) / p.k_per_split;
}

// one item per output tile, m-tiles in bands of `band` 128-row blocks (PRL_GROUP_M overrides) for L2 reuse of the B operand.
// Measured: 8 for the per-vocab-chunk GEMMs of the recompute backward (within 1 % of its neighbours); 4 (= one 512-row
// m-tile: neighbouring workers share the A panel) for the whole-vocabulary GEMMs of the saved-logits backward (+1..2 %).
static void schedule_tiles(GemmParams& p, int band = 8) {
  p.tiles_per_split = 1;
  p.n_splits = p.n_tiles;
  p.group_m = (int)(knob(K_GROUP_M) > 0 ? knob(K_GROUP_M) : band) / (cta_group() * msub(EPI_STORE));
  if (p.group_m < 1) p.group_m = 1;
  p.max_stages = (int)knob(K_TILE_STAGES);
  p.num_items = p.m_tiles * p.n_splits * (p.k_splits > 1 ? p.k_splits : 1);
}

// items = (m_tile, n-range): pick the number of vocab splits that minimises the static makespan
static void schedule_sweeps(GemmParams& p, int max_splits, int epi) {
  const int sms = workers();  // CTAs, CTA pairs or clusters of two pairs
  int best_ns = 1;
  long best = -1;
  const int hi = p.n_tiles < max_splits ? p.n_tiles : max_splits;
  for (int ns = 1; ns <= hi; ++ns) {
    const int tps = (p.n_tiles + ns - 1) / ns;
    const int ns_eff = (p.n_tiles + tps - 1) / tps;
    const long items = (long)p.m_tiles * ns_eff;
    const long makespan = ((items + sms - 1) / sms) * tps;
    if (best < 0 || makespan < best) {
      best = makespan;
      best_ns = ns_eff;
    }
  }
  p.tiles_per_split = (p.n_tiles + best_ns - 1) / best_ns;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.group_m = (int)knob(K_SWEEP_GROUP_M) / (cta_group() * msub(epi));
  if (p.group_m < 1) p.group_m = 1;
#ifdef PRL_DEBUG_KNOBS
  p.split_group = (int)knob(K_SPLIT_GROUP);
#endif
  p.num_items = p.m_tiles * p.n_splits;
}

constexpr int MAX_SPLITS = 64;

// ------------------------------------------------------------------------------------------------
// merge of per-split partials -> logp / entropy / lse (or one merged partial per token)
// ------------------------------------------------------------------------------------------------
__global__ void merge_partials_kernel(const float4* __restrict__ parts, int nparts, int64_t stride, int64_t T, int shift, float v_shard,
                                      float* __restrict__ logp, float* __restrict__ entropy, float* __restrict__ lse,
                                      float4* __restrict__ merged) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t - shift;
    float m, s, u, zt;
    if (r < 0) {
      // shift_logits' all-zero row (rl/loss.py:55-56): z == 0 over the whole shard
      m = 0.f, s = v_shard, u = 0.f, zt = 0.f;
    } else {
      m = -INFINITY, s = 0.f, u = 0.f, zt = 0.f;
      for (int i = 0; i < nparts; ++i) {
        const float4 q = parts[(int64_t)i * stride + r];
        const float mn = fmaxf(m, q.x);
        if (mn == -INFINITY) continue;
        const float ea = __expf(m - mn), eb = __expf(q.x - mn);
        s = s * ea + q.y * eb;
        u = u * ea + q.z * eb;
        zt += q.w;
        m = mn;
      }
    }
    if (merged) merged[t] = make_float4(m, s, u, zt);
    const float l = m + logf(s);
    if (lse) lse[t] = l;
    if (logp) logp[t] = zt - l;
    if (entropy) entropy[t] = l - u / s;
  }
}

// fp32 -> bf16 (RN) conversion of the d_hidden accumulator after the last vocab chunk
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n / 8;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n8; q += stride) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * q], b = reinterpret_cast<const float4*>(src)[2 * q + 1];
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y);
    w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y);
    w.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[q] = w;
  }
  for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

// Saved-logits backward, step 1: turn the bf16 logits the forward kept into dlogits IN PLACE (HBM-bound: one read and one
// write of [T, V] bf16).  Same per-element operations as the recompute epilogue (dz_values in gemm_engine.cuh), so both
// modes produce the same dlogits bit for bit.  One work item = 2048 consecutive columns of one row.
constexpr int DZS_THREADS = 256;
constexpr int DZS_SEG = DZS_THREADS * 8;

__device__ __forceinline__ uint4 dz_vec8(const uint4 in, float c1, float coef, float nlse2, int tj) {
  const uint64_t C1 = ptx::pack2(c1, c1), NL = ptx::pack2(nlse2, nlse2), NC = ptx::pack2(-coef, -coef);
  const uint32_t w[4] = {in.x, in.y, in.z, in.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a0 = __uint_as_float(w[k] << 16), a1 = __uint_as_float(w[k] & 0xFFFF0000u);
    float x0, x1, d0, d1;
    ptx::unpack2(ptx::fma2(ptx::pack2(a0, a1), C1, NL), x0, x1);
    ptx::unpack2(ptx::mul2(ptx::pack2(ptx::ex2(x0), ptx::ex2(x1)), NC), d0, d1);
    if (tj == 2 * k) d0 += coef;
    if (tj == 2 * k + 1) d1 += coef;
    o[k] = ptx::cvt_bf16x2(d0, d1);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(DZS_THREADS) dz_from_saved_logits_kernel(__nv_bfloat16* __restrict__ z, int64_t ld, int64_t T, int64_t V,
                                                                           int shift, const int64_t* __restrict__ ids,
                                                                           const float* __restrict__ lse, const float* __restrict__ dlogp,
                                                                           int64_t col_base, float inv_temp) {
  const int64_t n_seg = (V + DZS_SEG - 1) / DZS_SEG;
  const int64_t items = T * n_seg;
  const float c1 = inv_temp * LOG2E;
  const int64_t v8 = (V + 7) / 8 * 8;  // whole 16-byte vectors; columns in [V, v8) are padding of the caller's row pitch
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t row = item / n_seg;
    const int64_t col = (item - row * n_seg) * DZS_SEG + (int64_t)threadIdx.x * 8;
    if (col >= v8) continue;
    uint4* ptr = reinterpret_cast<uint4*>(z + row * ld + col);
    const int64_t t = row + shift;
    if (t >= T) {  // no target for this source row (the last row of a shifted packed row): zero gradient
      __stcs(ptr, make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    const uint4 in = __ldcs(ptr);
    const float coef = dlogp[t] * inv_temp;
    const float nlse2 = -lse[t] * LOG2E;
    const int64_t tj64 = ids[t] - col_base - col;
    const int tj = (tj64 >= 0 && tj64 < 8) ? (int)tj64 : -1;
    __stcs(ptr, dz_vec8(in, c1, coef, nlse2, tj));
  }
}

static int launch_merge(const float4* parts, int nparts, int64_t stride, int64_t T, int shift, float v_shard, float* logp, float* entropy,
                        float* lse, float4* merged, cudaStream_t st) {
  if (T <= 0) return PRL_OK;
  const int64_t need = (T + 255) / 256, cap = (int64_t)sm_count() * 4;
  merge_partials_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(parts, nparts, stride, T, shift, v_shard, logp, entropy, lse, merged);
  return check_launch("merge_partials_kernel");
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace prl

using namespace prl;

extern "C" {

size_t prl_lmhead_fwd_workspace_bytes(int64_t T, int32_t H, int64_t V) {
  (void)H;
  (void)V;
  return (size_t)MAX_SPLITS * (size_t)round_up(T > 0 ? T : 1, BLOCK_M) * sizeof(float4) + 256;
}

}  // extern "C"

// forward; `saved` != nullptr additionally keeps the bf16 logits ([T - shift rows][saved_ld]) for prl_lmhead_bwd_saved
static int lmhead_fwd_impl(const void* hidden, const void* weight, const int64_t* ids, int64_t T, int32_t H, int64_t V, int64_t v_offset,
                           int64_t V_total, float inv_temp, int32_t shift, int32_t round_bf16, float* logp, float* entropy, float* lse,
                           float* partials, void* saved, int64_t saved_ld, void* workspace, size_t workspace_bytes, void* stream) {
  PRL_REQUIRE(T >= 0 && H >= 1 && V >= 1, PRL_ERR_BAD_ARG, "lmhead_fwd: bad sizes");
  PRL_REQUIRE(shift == 0 || shift == 1, PRL_ERR_BAD_ARG, "lmhead_fwd: shift must be 0 or 1");
  PRL_REQUIRE(inv_temp > 0.f && inv_temp < INFINITY, PRL_ERR_BAD_ARG, "lmhead_fwd: temperature must be positive and finite");
  PRL_REQUIRE(v_offset >= 0 && v_offset + V <= V_total, PRL_ERR_BAD_ARG, "lmhead_fwd: vocab shard outside [0, V_total)");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(hidden && weight && ids && workspace, PRL_ERR_BAD_ARG, "lmhead_fwd: null pointer");
  PRL_REQUIRE(H % 8 == 0, PRL_ERR_UNSUPPORTED, "lmhead_fwd: H must be a multiple of 8 (16-byte TMA row pitch)");
  PRL_REQUIRE(workspace_bytes >= prl_lmhead_fwd_workspace_bytes(T, H, V), PRL_ERR_WORKSPACE, "lmhead_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = T - shift;  // source rows that feed an output position
  float4* parts = reinterpret_cast<float4*>(workspace);
  const int64_t stride = round_up(T, BLOCK_M);
  int nparts = 0;
  if (rows > 0) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_operand_map(&p.tma_a, hidden, false, rows, H, H, BLOCK_M);
    if (rc) return rc;
    rc = make_operand_map(&p.tma_b, weight, false, V, H, H, BLOCK_N);
    if (rc) return rc;
    set_tiling(p, rows, V, H, EPI_STATS);
    schedule_sweeps(p, MAX_SPLITS, EPI_STATS);
    make_desc_defaults(p, false, false);
    p.epi.m_rows = rows;
    p.epi.n_cols = V;
    p.epi.ids = ids;
    p.epi.T = T;
    p.epi.shift = shift;
    p.epi.round_bf16 = round_bf16;
    p.epi.col_base = v_offset;
    p.epi.inv_temp = inv_temp;
    p.epi.partials = parts;
    p.epi.part_stride = stride;
    // Prefetch depth: 5 of the 7 stages that fit.  Measured (ncu, one forward at the Qwen3-8B shape): 7 stages 18-23 GB of DRAM
    // reads, 5 stages 11.6 GB, 4 stages 6.7 GB (3 stages starve the MMAs); sustained under the power cap 5 stages is fastest
    // (+4..6 % over 7).  The hidden band is marked evict-last.
    p.max_stages = (int)knob(K_FWD_STAGES);
    if (saved) {
      p.epi.dz = reinterpret_cast<__nv_bfloat16*>(saved);
      p.epi.ldz = saved_ld;
      if (epi_tma_enabled()) {
        rc = make_output_map(&p.tma_c, saved, rows, V, saved_ld, false);
        if (rc) return rc;
        p.epi.flags = ST_TMA;
      }
      set_hints(p, "FWDS", 2, 0, 1);  // the kept logits are not read again before the backward: evict first
      rc = launch_gemm<false, false, EPI_STATS_STORE>(p, st, "lmhead_fwd_save_kernel");
    } else {
      set_hints(p, "FWD", 2, 0, 0);
      rc = launch_gemm<false, false, EPI_STATS>(p, st, "lmhead_fwd_kernel");
    }
    if (rc) return rc;
    nparts = p.n_splits;
  }
  return launch_merge(parts, nparts, stride, T, shift, (float)V, logp, entropy, lse, reinterpret_cast<float4*>(partials), st);
}

extern "C" {

int prl_lmhead_fwd(const void* hidden, const void* weight, const int64_t* ids, int64_t T, int32_t H, int64_t V, int64_t v_offset,
                   int64_t V_total, float inv_temp, int32_t shift, int32_t round_bf16, float* logp, float* entropy, float* lse,
                   float* partials, void* workspace, size_t workspace_bytes, void* stream) {
  return lmhead_fwd_impl(hidden, weight, ids, T, H, V, v_offset, V_total, inv_temp, shift, round_bf16, logp, entropy, lse, partials,
                         nullptr, 0, workspace, workspace_bytes, stream);
}

int64_t prl_lmhead_saved_logits_ld(int64_t V) { return round_up(V > 0 ? V : 1, 64); }

int prl_lmhead_fwd_save(const void* hidden, const void* weight, const int64_t* ids, int64_t T, int32_t H, int64_t V, int64_t v_offset,
                        int64_t V_total, float inv_temp, int32_t shift, float* logp, float* entropy, float* lse, float* partials,
                        void* saved_logits, int64_t saved_ld, void* workspace, size_t workspace_bytes, void* stream) {
  PRL_REQUIRE(T == 0 || saved_logits, PRL_ERR_BAD_ARG, "lmhead_fwd_save: null saved_logits");
  PRL_REQUIRE(saved_ld >= V && saved_ld % 8 == 0, PRL_ERR_BAD_ARG, "lmhead_fwd_save: saved_ld must be >= V and a multiple of 8");
  PRL_REQUIRE((reinterpret_cast<uintptr_t>(saved_logits) & 15) == 0, PRL_ERR_BAD_ARG, "lmhead_fwd_save: saved_logits must be 16-byte aligned");
  return lmhead_fwd_impl(hidden, weight, ids, T, H, V, v_offset, V_total, inv_temp, shift, 1, logp, entropy, lse, partials, saved_logits,
                         saved_ld, workspace, workspace_bytes, stream);
}

int prl_lmhead_merge_partials(const float* partials, int32_t nparts, int64_t T, float* logp, float* entropy, float* lse, void* stream) {
  PRL_REQUIRE(nparts >= 1 && T >= 0, PRL_ERR_BAD_ARG, "merge_partials: bad sizes");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(partials, PRL_ERR_BAD_ARG, "merge_partials: null pointer");
  PRL_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 15) == 0, PRL_ERR_BAD_ARG, "merge_partials: partials must be 16-byte aligned");
  return launch_merge(reinterpret_cast<const float4*>(partials), nparts, T, T, 0, 0.f, logp, entropy, lse, nullptr, (cudaStream_t)stream);
}

}  // extern "C"

// fused reduce-scatter of d_hidden (vocab-sharded LM head): rows [j * rows_per_peer, (j + 1) * rows_per_peer) go to ptrs[j]
struct PeerScatter {
  void* const* ptrs;      // host array of n DEVICE pointers: every peer's fp32 [rows_per_peer, H] accumulator (own one included)
  int32_t n;
  int64_t rows_per_peer;
  int32_t self;           // this rank's index (sets the m-tile / vocab-chunk rotation)
  int64_t total_rows;     // d_weight scatter: rows of the whole matrix (the last owners may hold fewer than rows_per_peer)
  const float* local_add; // d_weight scatter: this rank's local fp32 [total_rows, H] accumulator to add on the way out (or NULL)
};

// Fused reduce-scatter: rows [j * rows_per_peer, ..) of the output belong to peer j; rank `self` starts at the tiles of the NEXT
// rank's tokens, so that at any time the ranks (which walk their tile lists at the same pace) write to different owners.
static void set_scatter_rotation(GemmParams& p, int n_peers, int64_t rows_per_peer, int self) {
  p.n_peers = n_peers;
  p.peer_rows = (int32_t)rows_per_peer;
  const int64_t rows_per_tile = (int64_t)BLOCK_M * cta_group() * msub(EPI_STORE);
  p.m_rot = (int32_t)(((((int64_t)self + 1) % n_peers) * rows_per_peer / rows_per_tile) % p.m_tiles);
}

// d_hidden and d_weight GEMMs over one block of dlogits: dz[T][ldz] holds vocab columns [c0, c0 + vce) of this shard
static int lmhead_bwd_gemms(const void* hidden, const __nv_bfloat16* W, const __nv_bfloat16* dz, int64_t vc, int64_t c0, int64_t vce,
                            int64_t T, int32_t H, void* d_hidden, int32_t d_hidden_f32, void* d_weight, int32_t d_weight_dtype,
                            int32_t accumulate_dw, int32_t phase_mask, bool first, bool last, int64_t nchunks, float* dh_acc,
                            int band, cudaStream_t st, const PeerScatter* peers = nullptr, const PeerScatter* dw_peers = nullptr) {
  const int64_t vc_full = nchunks > 1 ? vc : vce;  // every chunk of one backward takes the same k-split decision
  // ---- K2: d_hidden[T,H] += dz[T,vce] x W[c0:c0+vce, :]   (A K-major, B MN-major) ----
  if (peers && (phase_mask & 2)) {
    // every tile is reduce-added into its owner's accumulator (zeroed by the owner before the round): no local d_hidden buffer
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_operand_map(&p.tma_a, dz, false, T, vce, vc, BLOCK_M);
    if (rc) return rc;
    rc = make_operand_map(&p.tma_b, W + c0 * H, true, H, vce, H, BLOCK_N);
    if (rc) return rc;
    set_tiling(p, T, H, vce, EPI_STORE);
    schedule_tiles(p, band);
    make_desc_defaults(p, false, true);
    p.epi.m_rows = T;
    p.epi.n_cols = H;
    p.epi.ldc = H;
    p.epi.flags = ST_STORE_F32 | ST_LOAD_F32 | ST_TMA;
    for (int j = 0; j < peers->n; ++j) {
      rc = make_output_map(&p.tma_c_peer[j], peers->ptrs[j], peers->rows_per_peer, H, H, true);
      if (rc) return rc;
      p.peer_nrows[j] = (int32_t)peers->rows_per_peer;
      p.peer_ptr[j] = reinterpret_cast<float*>(peers->ptrs[j]);
    }
    set_scatter_rotation(p, peers->n, peers->rows_per_peer, peers->self);
    set_hints(p, "DH", 0, 0, 0);
    rc = launch_gemm<false, true, EPI_STORE>(p, st, "lmhead_bwd_dhidden_rs_kernel");
    if (rc) return rc;
  } else if (d_hidden && (phase_mask & 2)) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_operand_map(&p.tma_a, dz, false, T, vce, vc, BLOCK_M);
    if (rc) return rc;
    rc = make_operand_map(&p.tma_b, W + c0 * H, true, H, vce, H, BLOCK_N);
    if (rc) return rc;
    set_tiling(p, T, H, vce, EPI_STORE);
    const bool tma = epi_tma_enabled();
    // k-splits need an fp32 accumulator that is reduced into: d_hidden itself (fp32 output) or dh_acc
    const int ks = (tma && (d_hidden_f32 || dh_acc)) ? pick_k_splits(T, H, vc_full) : 1;
    set_k_splits(p, ks);
    const bool split = p.k_splits > 1;
    schedule_tiles(p, band);
    make_desc_defaults(p, false, true);
    p.epi.m_rows = T;
    p.epi.n_cols = H;
    p.epi.ldc = H;
    if (split && first) {  // every item of every chunk reduce-adds into a zeroed accumulator
      cudaError_t e = cudaMemsetAsync(d_hidden_f32 ? d_hidden : (void*)dh_acc, 0, (size_t)T * H * sizeof(float), st);
      if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(d_hidden accumulator)");
    }
    if (split) first = false;
    if (d_hidden_f32) {
      p.epi.c_f32 = reinterpret_cast<float*>(d_hidden);
      p.epi.flags = ST_STORE_F32 | (first ? 0 : ST_LOAD_F32);
      if (tma) {
        rc = make_output_map(&p.tma_c, d_hidden, T, H, H, true);
        p.epi.flags |= ST_TMA;
      }
    } else if (tma && nchunks == 1 && !split) {
      p.epi.c_bf16 = reinterpret_cast<__nv_bfloat16*>(d_hidden);
      p.epi.flags = ST_STORE_BF16 | ST_TMA;
      rc = make_output_map(&p.tma_c, d_hidden, T, H, H, false);
    } else if (tma) {
      // fp32 accumulator in the workspace: TMA store for the first vocab chunk, TMA reduce-add (performed in L2, no
      // SM-side read) for the others; one conversion pass to bf16 after the last chunk
      p.epi.c_f32 = dh_acc;
      p.epi.flags = ST_STORE_F32 | (first ? 0 : ST_LOAD_F32) | ST_TMA;
      rc = make_output_map(&p.tma_c, dh_acc, T, H, H, true);
    } else {
      p.epi.c_f32 = dh_acc;
      p.epi.c_bf16 = reinterpret_cast<__nv_bfloat16*>(d_hidden);
      p.epi.flags = (first ? 0 : ST_LOAD_F32) | (last ? ST_STORE_BF16 : ST_STORE_F32);
    }
    if (rc) return rc;
    set_hints(p, "DH", 0, 0, 0);
    rc = launch_gemm<false, true, EPI_STORE>(p, st, "lmhead_bwd_dhidden_kernel");
    if (rc) return rc;
    if (last && tma && !d_hidden_f32 && (nchunks > 1 || split)) {
      const int64_t n = T * (int64_t)H;
      const int64_t need = (n / 8 + 255) / 256 + 1, cap = (int64_t)sm_count() * 8;
      f32_to_bf16_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(dh_acc, reinterpret_cast<__nv_bfloat16*>(d_hidden), n);
      rc = check_launch("f32_to_bf16_kernel");
      if (rc) return rc;
    }
  }
  // ---- K3: d_weight[c0:c0+vce, H] (+)= dz^T x hidden   (both MN-major) ----
  if ((d_weight || dw_peers) && (phase_mask & 4)) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_operand_map(&p.tma_a, dz, true, vce, T, vc, BLOCK_M);
    if (rc) return rc;
    rc = make_operand_map(&p.tma_b, hidden, true, H, T, H, BLOCK_N);
    if (rc) return rc;
    set_tiling(p, vce, H, T, EPI_STORE);
    schedule_tiles(p, band);
    make_desc_defaults(p, true, true);
    p.epi.m_rows = vce;
    p.epi.n_cols = H;
    p.epi.ldc = H;
    if (dw_peers) {
      // gradient accumulation straight into the reduce-scattered layout: the rows of this chunk are reduce-added (fp32) into the
      // shards of the ranks that own them; a chunk spans at most a few owners
      p.epi.flags = ST_STORE_F32 | ST_LOAD_F32 | ST_TMA;
      if (dw_peers->local_add) {  // send (local accumulator + this micro-batch's tile): the step's total leaves in ONE pass
        p.epi.flags |= ST_ADD_LOCAL;
        p.epi.c_f32 = const_cast<float*>(dw_peers->local_add) + c0 * H;
      }
      p.n_peers = dw_peers->n;
      p.peer_rows = (int32_t)dw_peers->rows_per_peer;
      p.peer_row0 = (int32_t)c0;
      for (int j = 0; j < dw_peers->n; ++j) {
        const int64_t rows_j = dw_peers->total_rows - (int64_t)j * dw_peers->rows_per_peer;
        const int64_t rows_here = rows_j < dw_peers->rows_per_peer ? rows_j : dw_peers->rows_per_peer;
        if (rows_here <= 0) {  // FSDP's trailing ranks may own nothing (torch.chunk semantics)
          p.n_peers = j;
          break;
        }
        rc = make_output_map(&p.tma_c_peer[j], dw_peers->ptrs[j], rows_here, H, H, true);
        if (rc) return rc;
        p.peer_nrows[j] = (int32_t)rows_here;
        p.peer_ptr[j] = reinterpret_cast<float*>(dw_peers->ptrs[j]);
      }
    } else if (d_weight_dtype == PRL_DT_F32) {
      p.epi.c_f32 = reinterpret_cast<float*>(d_weight) + c0 * H;
      p.epi.flags = ST_STORE_F32 | (accumulate_dw ? ST_LOAD_F32 : 0);
      if (epi_tma_enabled()) {
        rc = make_output_map(&p.tma_c, p.epi.c_f32, vce, H, H, true);
        if (rc) return rc;
        p.epi.flags |= ST_TMA;
      }
    } else {
      p.epi.c_bf16 = reinterpret_cast<__nv_bfloat16*>(d_weight) + c0 * H;
      p.epi.flags = ST_STORE_BF16 | (accumulate_dw ? ST_LOAD_BF16 : 0);
      if (epi_tma_enabled() && !accumulate_dw) {
        rc = make_output_map(&p.tma_c, p.epi.c_bf16, vce, H, H, false);
        if (rc) return rc;
        p.epi.flags |= ST_TMA;
      }
    }
    set_hints(p, "DW", 0, 0, 0);
    rc = launch_gemm<true, true, EPI_STORE>(p, st, "lmhead_bwd_dweight_kernel");
    if (rc) return rc;
  }
  return PRL_OK;
}

extern "C" {

static int64_t pick_vchunk(int64_t V, int64_t vchunk) {
  if (vchunk <= 0) vchunk = knob(K_VCHUNK);
  vchunk = round_up(vchunk, BLOCK_N);
  const int64_t vpad = round_up(V, BLOCK_N);
  return vchunk < vpad ? vchunk : vpad;
}

size_t prl_lmhead_bwd_workspace_bytes(int64_t T, int32_t H, int64_t V, int64_t vchunk) {
  const int64_t vc = pick_vchunk(V, vchunk);
  const size_t dz = (size_t)round_up(T > 0 ? T : 1, BLOCK_M) * (size_t)vc * 2;
  const size_t dh = (size_t)(T > 0 ? T : 1) * (size_t)H * 4;
  return round_up((int64_t)dz, 1024) + round_up((int64_t)dh, 1024) + 1024;
}

static int lmhead_bwd_impl(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp, int64_t T,
                           int32_t H, int64_t V, int64_t v_offset, float inv_temp, int32_t shift, int32_t round_bf16, void* d_hidden,
                           int32_t d_hidden_f32, void* d_weight, int32_t d_weight_dtype, int32_t accumulate_dw, int64_t vchunk,
                           int32_t phase_mask, void* workspace, size_t workspace_bytes, void* stream, const PeerScatter* peers,
                           const PeerScatter* dw_peers = nullptr) {
  PRL_REQUIRE(T >= 0 && H >= 1 && V >= 1, PRL_ERR_BAD_ARG, "lmhead_bwd: bad sizes");
  PRL_REQUIRE(shift == 0 || shift == 1, PRL_ERR_BAD_ARG, "lmhead_bwd: shift must be 0 or 1");
  PRL_REQUIRE(inv_temp > 0.f && inv_temp < INFINITY, PRL_ERR_BAD_ARG, "lmhead_bwd: temperature must be positive and finite");
  PRL_REQUIRE(d_weight_dtype == PRL_DT_BF16 || d_weight_dtype == PRL_DT_F32, PRL_ERR_BAD_ARG, "lmhead_bwd: bad d_weight dtype");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(hidden && weight && ids && lse && dlogp && workspace, PRL_ERR_BAD_ARG, "lmhead_bwd: null pointer");
  PRL_REQUIRE(H % 8 == 0, PRL_ERR_UNSUPPORTED, "lmhead_bwd: H must be a multiple of 8 (16-byte TMA row pitch)");
  PRL_REQUIRE(workspace_bytes >= prl_lmhead_bwd_workspace_bytes(T, H, V, vchunk), PRL_ERR_WORKSPACE, "lmhead_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t vc = pick_vchunk(V, vchunk);
  const int64_t nchunks = (V + vc - 1) / vc;
  // d_weight scatter: every rank starts at a different vocab chunk, so that at any time the ranks add into different owners' shards
  const int64_t chunk_rot = dw_peers ? ((int64_t)dw_peers->self * nchunks) / dw_peers->n : 0;
  const int64_t t_pad = round_up(T, BLOCK_M);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(workspace);
  float* dh_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + round_up(t_pad * vc * 2, 1024));
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(weight);

  if ((phase_mask & 7) == 0) phase_mask = 7;
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t c = (ci + chunk_rot) % nchunks;
    const int64_t c0 = c * vc;
    const int64_t vce = (V - c0) < vc ? (V - c0) : vc;
    // ---- K1: recompute the logits tile by tile, emit dlogits (bf16) for this vocab chunk ----
    if (phase_mask & 1) {
      GemmParams p;
      memset(&p, 0, sizeof(p));
      int rc = make_operand_map(&p.tma_a, hidden, false, T, H, H, BLOCK_M);
      if (rc) return rc;
      rc = make_operand_map(&p.tma_b, W + c0 * H, false, vce, H, H, BLOCK_N);
      if (rc) return rc;
      set_tiling(p, T, vce, H, EPI_DZ);
      schedule_sweeps(p, MAX_SPLITS, EPI_DZ);
      make_desc_defaults(p, false, false);
      p.epi.m_rows = T;
      p.epi.n_cols = vce;
      p.epi.ids = ids;
      p.epi.T = T;
      p.epi.shift = shift;
      p.epi.round_bf16 = round_bf16;
      p.epi.col_base = v_offset + c0;
      p.epi.inv_temp = inv_temp;
      p.epi.lse = lse;
      p.epi.dlogp = dlogp;
      p.epi.dz = dz;
      p.epi.ldz = vc;
      if (epi_tma_enabled()) {
        rc = make_output_map(&p.tma_c, dz, T, vce, vc, false);
        if (rc) return rc;
        p.epi.flags = ST_TMA;
      }
      set_hints(p, "DZ", 2, 1, 1);  // hidden band reused across the sweep, W tiles and the dlogits block stream through
      p.max_stages = (int)knob(K_DZ_STAGES);  // 5 of 6 stages: +1 % over 6, 4 is 6 % slower
      rc = launch_gemm<false, false, EPI_DZ>(p, st, "lmhead_bwd_dz_kernel");
      if (rc) return rc;
    }
    {
      const int rc = lmhead_bwd_gemms(hidden, W, dz, vc, c0, vce, T, H, d_hidden, d_hidden_f32, d_weight, d_weight_dtype, accumulate_dw,
                                      phase_mask, ci == 0, ci == nchunks - 1, nchunks, dh_acc, 8, st, peers, dw_peers);
      if (rc) return rc;
    }
  }
  return PRL_OK;
}

int prl_lmhead_bwd(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp, int64_t T, int32_t H,
                   int64_t V, int64_t v_offset, float inv_temp, int32_t shift, int32_t round_bf16, void* d_hidden, int32_t d_hidden_f32,
                   void* d_weight, int32_t d_weight_dtype, int32_t accumulate_dw, int64_t vchunk, int32_t phase_mask, void* workspace,
                   size_t workspace_bytes, void* stream) {
  return lmhead_bwd_impl(hidden, weight, ids, lse, dlogp, T, H, V, v_offset, inv_temp, shift, round_bf16, d_hidden, d_hidden_f32, d_weight,
                         d_weight_dtype, accumulate_dw, vchunk, phase_mask, workspace, workspace_bytes, stream, nullptr);
}

int prl_lmhead_bwd_scatter(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp, int64_t T,
                           int32_t H, int64_t V, int64_t v_offset, float inv_temp, int32_t shift, int32_t round_bf16,
                           void* const* d_hidden_peers, int32_t n_peers, int64_t rows_per_peer, int32_t self_index, void* d_weight,
                           int32_t d_weight_dtype, int32_t accumulate_dw, int64_t vchunk, int32_t phase_mask, void* workspace,
                           size_t workspace_bytes, void* stream) {
  PRL_REQUIRE(d_hidden_peers && n_peers >= 1 && n_peers <= 8 && self_index >= 0 && self_index < n_peers, PRL_ERR_BAD_ARG,
              "lmhead_bwd_scatter: 1..8 peers");
  PRL_REQUIRE(rows_per_peer >= 32 && rows_per_peer % 32 == 0 && rows_per_peer < ((int64_t)1 << 31), PRL_ERR_UNSUPPORTED,
              "lmhead_bwd_scatter: rows_per_peer must be a multiple of 32 (one epilogue slab never straddles two owners)");
  PRL_REQUIRE(T == rows_per_peer * n_peers, PRL_ERR_BAD_ARG, "lmhead_bwd_scatter: T must be n_peers * rows_per_peer");
  PRL_REQUIRE(epi_tma_enabled(), PRL_ERR_UNSUPPORTED, "lmhead_bwd_scatter needs the TMA-store epilogue (PRL_EPI_TMA=1)");
  for (int j = 0; j < n_peers; ++j) PRL_REQUIRE(d_hidden_peers[j], PRL_ERR_BAD_ARG, "lmhead_bwd_scatter: null peer pointer");
  PeerScatter ps = {d_hidden_peers, n_peers, rows_per_peer, self_index, T, nullptr};
  return lmhead_bwd_impl(hidden, weight, ids, lse, dlogp, T, H, V, v_offset, inv_temp, shift, round_bf16, nullptr, 1, d_weight, d_weight_dtype,
                         accumulate_dw, vchunk, phase_mask, workspace, workspace_bytes, stream, &ps);
}

int prl_lmhead_bwd_dw_scatter(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp, int64_t T,
                              int32_t H, int64_t V, float inv_temp, int32_t shift, int32_t round_bf16, void* d_hidden, int32_t d_hidden_f32,
                              void* const* d_weight_shards, int32_t n_peers, int64_t rows_per_peer, int32_t self_index,
                              const float* local_accumulator, int64_t vchunk, int32_t phase_mask, void* workspace, size_t workspace_bytes,
                              void* stream) {
  PRL_REQUIRE(d_weight_shards && n_peers >= 1 && n_peers <= 8 && self_index >= 0 && self_index < n_peers, PRL_ERR_BAD_ARG,
              "lmhead_bwd_dw_scatter: 1..8 peers");
  PRL_REQUIRE(rows_per_peer >= 32 && rows_per_peer < ((int64_t)1 << 31) && rows_per_peer * n_peers >= V && V < ((int64_t)1 << 31),
              PRL_ERR_BAD_ARG, "lmhead_bwd_dw_scatter: rows_per_peer must be ceil(V / n_peers) (>= 32)");
  PRL_REQUIRE(epi_tma_enabled(), PRL_ERR_UNSUPPORTED, "lmhead_bwd_dw_scatter needs the TMA-store epilogue (PRL_EPI_TMA=1)");
  for (int j = 0; j < n_peers; ++j)
    PRL_REQUIRE(d_weight_shards[j] || (int64_t)j * rows_per_peer >= V, PRL_ERR_BAD_ARG, "lmhead_bwd_dw_scatter: null shard pointer");
  PRL_REQUIRE((reinterpret_cast<uintptr_t>(local_accumulator) & 15) == 0, PRL_ERR_BAD_ARG, "lmhead_bwd_dw_scatter: local accumulator must be 16-byte aligned");
  PeerScatter ps = {d_weight_shards, n_peers, rows_per_peer, self_index, V, local_accumulator};
  return lmhead_bwd_impl(hidden, weight, ids, lse, dlogp, T, H, V, 0, inv_temp, shift, round_bf16, d_hidden, d_hidden_f32, nullptr, PRL_DT_F32, 1,
                         vchunk, phase_mask, workspace, workspace_bytes, stream, nullptr, &ps);
}

size_t prl_lmhead_bwd_saved_workspace_bytes(int64_t T, int32_t H, int64_t V) {
  // an fp32 [T, H] accumulator, needed only when the d_hidden GEMM is split along K (few output tiles: small T x H)
  if (T <= 0 || H <= 0 || V <= 0 || pick_k_splits(T, H, V) <= 1) return 0;
  return (size_t)T * (size_t)H * sizeof(float) + 1024;
}

int prl_lmhead_bwd_saved(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp, int64_t T,
                         int32_t H, int64_t V, int64_t v_offset, float inv_temp, int32_t shift, void* saved_logits, int64_t saved_ld,
                         void* d_hidden, int32_t d_hidden_f32, void* d_weight, int32_t d_weight_dtype, int32_t accumulate_dw,
                         int32_t phase_mask, void* workspace, size_t workspace_bytes, void* stream) {
  PRL_REQUIRE(T >= 0 && H >= 1 && V >= 1, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: bad sizes");
  PRL_REQUIRE(shift == 0 || shift == 1, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: shift must be 0 or 1");
  PRL_REQUIRE(inv_temp > 0.f && inv_temp < INFINITY, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: temperature must be positive and finite");
  PRL_REQUIRE(d_weight_dtype == PRL_DT_BF16 || d_weight_dtype == PRL_DT_F32, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: bad d_weight dtype");
  PRL_REQUIRE(saved_ld >= V && saved_ld % 8 == 0, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: saved_ld must be >= V and a multiple of 8");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(hidden && weight && ids && lse && dlogp && saved_logits, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: null pointer");
  PRL_REQUIRE((reinterpret_cast<uintptr_t>(saved_logits) & 15) == 0, PRL_ERR_BAD_ARG, "lmhead_bwd_saved: saved_logits must be 16-byte aligned");
  PRL_REQUIRE(H % 8 == 0, PRL_ERR_UNSUPPORTED, "lmhead_bwd_saved: H must be a multiple of 8 (16-byte TMA row pitch)");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(saved_logits);
  if ((phase_mask & 7) == 0) phase_mask = 7;
  if (phase_mask & 1) {
    const int64_t items = T * ((V + DZS_SEG - 1) / DZS_SEG), cap = (int64_t)sm_count() * 16;
    dz_from_saved_logits_kernel<<<(int)(items < cap ? items : cap), DZS_THREADS, 0, st>>>(dz, saved_ld, T, V, shift, ids, lse, dlogp, v_offset,
                                                                                         inv_temp);
    const int rc = check_launch("dz_from_saved_logits_kernel");
    if (rc) return rc;
  }
  // optional fp32 [T, H] accumulator (1024-byte aligned inside the workspace); without it the d_hidden GEMM is not k-split
  float* dh_acc = nullptr;
  if (workspace && workspace_bytes >= (size_t)T * H * sizeof(float) + 1024)
    dh_acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  return lmhead_bwd_gemms(hidden, reinterpret_cast<const __nv_bfloat16*>(weight), dz, saved_ld, 0, V, T, H, d_hidden, d_hidden_f32, d_weight,
                          d_weight_dtype, accumulate_dw, phase_mask, true, true, 1, dh_acc, 4, st);
}

int prl_set_knob(const char* name, int64_t value) {
  PRL_REQUIRE(name, PRL_ERR_BAD_ARG, "set_knob: null name");
  KnobTable& t = knobs();
  for (int i = 0; i < K_COUNT; ++i)
    if (strcmp(t.names[i], name) == 0) {
      t.values[i] = (long)value;
      return PRL_OK;
    }
  return fail(PRL_ERR_BAD_ARG, "set_knob: unknown knob %s (debug knobs need a -DPRL_DEBUG_KNOBS build)", name);
}

int prl_get_knob(const char* name, int64_t* value) {
  PRL_REQUIRE(name && value, PRL_ERR_BAD_ARG, "get_knob: null pointer");
  KnobTable& t = knobs();
  for (int i = 0; i < K_COUNT; ++i)
    if (strcmp(t.names[i], name) == 0) {
      *value = t.values[i];
      return PRL_OK;
    }
  return fail(PRL_ERR_BAD_ARG, "get_knob: unknown knob %s", name);
}

int prl_debug_schedule(int32_t kind, int64_t M, int64_t N, int64_t K, int32_t* items, int64_t capacity, int32_t* dims) {
  PRL_REQUIRE(kind >= 0 && kind <= 3 && M >= 1 && N >= 1 && K >= 1 && dims, PRL_ERR_BAD_ARG, "debug_schedule: bad arguments");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int epi = kind == 0 ? EPI_STATS : (kind == 1 ? EPI_DZ : EPI_STORE);
  set_tiling(p, M, N, K, epi);
  if (kind <= 1) {
    schedule_sweeps(p, MAX_SPLITS, epi);
  } else {
    set_k_splits(p, pick_k_splits(M, N, K));
    schedule_tiles(p, kind == 2 ? 8 : 4);
  }
  dims[0] = p.m_tiles, dims[1] = p.n_tiles, dims[2] = p.k_blocks, dims[3] = p.num_items;
  dims[4] = workers(), dims[5] = p.k_splits, dims[6] = p.n_splits, dims[7] = BLOCK_M * cta_group() * msub(epi);
  if (items) {
    for (int i = 0; i < p.num_items && i < capacity; ++i) {
      const Item it = decode_item(p, i);
      int32_t* o = items + 6 * (int64_t)i;
      o[0] = it.m_tile, o[1] = it.n_begin, o[2] = it.n_count, o[3] = it.k_begin, o[4] = it.k_count, o[5] = it.split;
    }
  }
  return PRL_OK;
}

int prl_debug_schedule_scatter(int64_t M, int64_t N, int64_t K, int32_t n_peers, int32_t self_index, int32_t* items, int64_t capacity,
                               int32_t* dims) {
  PRL_REQUIRE(M >= 1 && N >= 1 && K >= 1 && dims && n_peers >= 1 && n_peers <= 8 && self_index >= 0 && self_index < n_peers && M % n_peers == 0,
              PRL_ERR_BAD_ARG, "debug_schedule_scatter: bad arguments");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  set_tiling(p, M, N, K, EPI_STORE);
  schedule_tiles(p, 8);
  set_scatter_rotation(p, n_peers, M / n_peers, self_index);
  dims[0] = p.m_tiles, dims[1] = p.n_tiles, dims[2] = p.k_blocks, dims[3] = p.num_items;
  dims[4] = workers(), dims[5] = p.m_rot, dims[6] = p.peer_rows, dims[7] = BLOCK_M * cta_group() * msub(EPI_STORE);
  if (items) {
    for (int i = 0; i < p.num_items && i < capacity; ++i) {
      const Item it = decode_item(p, i);
      int32_t* o = items + 6 * (int64_t)i;
      o[0] = it.m_tile, o[1] = it.n_begin, o[2] = it.n_count, o[3] = it.k_begin, o[4] = it.k_count, o[5] = it.split;
    }
  }
  return PRL_OK;
}

int prl_gemm_bf16_f32(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int32_t a_mn_major, int32_t b_mn_major,
                      int32_t accumulate, void* stream) {
  PRL_REQUIRE(M >= 1 && N >= 1 && K >= 1, PRL_ERR_BAD_ARG, "gemm: bad sizes");
  PRL_REQUIRE(A && B && C, PRL_ERR_BAD_ARG, "gemm: null pointer");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_operand_map(&p.tma_a, A, a_mn_major != 0, M, K, a_mn_major ? M : K, BLOCK_M);
  if (rc) return rc;
  rc = make_operand_map(&p.tma_b, B, b_mn_major != 0, N, K, b_mn_major ? N : K, BLOCK_N);
  if (rc) return rc;
  set_tiling(p, M, N, K, EPI_STORE);
  schedule_tiles(p);
  make_desc_defaults(p, a_mn_major != 0, b_mn_major != 0);
  p.epi.m_rows = M;
  p.epi.n_cols = N;
  p.epi.ldc = N;
  p.epi.c_f32 = C;
  p.epi.flags = ST_STORE_F32 | (accumulate ? ST_LOAD_F32 : 0);
  if (epi_tma_enabled()) {
    rc = make_output_map(&p.tma_c, C, M, N, N, true);
    if (rc) return rc;
    p.epi.flags |= ST_TMA;
  }
  set_hints(p, "GEMM", 0, 0, 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (!a_mn_major && !b_mn_major) return launch_gemm<false, false, EPI_STORE>(p, st, "gemm_kk");
  if (!a_mn_major && b_mn_major) return launch_gemm<false, true, EPI_STORE>(p, st, "gemm_km");
  if (a_mn_major && !b_mn_major) return launch_gemm<true, false, EPI_STORE>(p, st, "gemm_mk");
  return launch_gemm<true, true, EPI_STORE>(p, st, "gemm_mm");
}

}  // extern "C"
