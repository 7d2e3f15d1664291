// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a with pluggable epilogues.
//
//   C[M,N] = A x B, bf16 inputs, fp32 accumulation in TMEM, BLOCK_K = 64.  Tile shapes (template parameters CG, MS):
//     CG = 1         128 x 256 per CTA
//     CG = 2, MS = 1 256 x 256 per CTA pair (tcgen05.mma.cta_group::2), double-buffered accumulators   [forward, dlogits]
//     CG = 2, MS = 2 512 x 256 per CTA pair: two M sub-tiles share every B k-block                     [d_hidden, d_weight]
//     CG = 4         two pairs per cluster sharing B by TMA multicast (kept for comparison)
//   warp 0 (one lane): TMA producer  -> smem ring of 128B-swizzled stages (as many as fit: 3..7; the host may cap the depth)
//   warp 1 (one lane): tcgen05.mma issuer -> two 256-column TMEM accumulators
//   warp 2           : TMEM allocator
//   warps 4..        : epilogue (4 warps per M sub-tile); thread i owns accumulator row (lane) i of the tile, so row-wise
//                      softmax statistics are thread-local (tcgen05.ld 32x32b); outputs leave through smem staging + TMA
//                      store / reduce-add.
// Each operand is either K-major (rows of K contiguous) or MN-major (K rows of contiguous M/N), both in
// the canonical SWIZZLE_128B layouts written directly by TMA.
//
// A work item is (m_tile, contiguous range of n_tiles): the epilogue keeps per-row online state across
// the n-range (forward statistics) and flushes at the end of the item.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace prl {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
#ifndef PRL_STAGES
#define PRL_STAGES 4
#endif
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

// EPI_STATS_STORE = the forward statistics plus a bf16 copy of the (rounded) logits, kept for the saved-logits backward
enum EpiKind { EPI_STATS = 0, EPI_DZ = 1, EPI_STORE = 2, EPI_STATS_STORE = 3 };

// flags of the STORE epilogue
constexpr int ST_LOAD_F32 = 1;    // C += existing fp32 at c_f32
constexpr int ST_STORE_F32 = 2;   // write fp32 to c_f32
constexpr int ST_STORE_BF16 = 4;  // write bf16 to c_bf16
constexpr int ST_LOAD_BF16 = 8;   // C += existing bf16 at c_bf16
constexpr int ST_TMA = 16;         // epilogue writes through smem staging + TMA store / reduce-add (tma_c)
constexpr int ST_ADD_LOCAL = 32;   // fused scatter: add the caller's LOCAL fp32 accumulator tile (c_f32, ldc) to the tile before it
                                   // leaves for its owner (the last micro-batch of a step sends local total + its own contribution)
// Profiling aids, compiled in only with -DPRL_DEBUG_KNOBS (they give WRONG results by design); in the release build the flags are
// 0 and every branch on them folds away.
#ifdef PRL_DEBUG_KNOBS
constexpr int ST_WAIT_BACKOFF = 1 << 18;  // epilogue warps sleep between mbarrier probes (PRL_WAIT_BACKOFF=1)
constexpr int ST_DEBUG_SKIPB = 1 << 17;   // skip the B load on odd k-blocks (PRL_DEBUG_SKIPB=1): cost of the L2->SM operand traffic
constexpr int ST_DEBUG_SKIP = 1 << 16;    // run the GEMM but skip the epilogue's global memory traffic (PRL_DEBUG_NOSTORE=1)
#else
constexpr int ST_WAIT_BACKOFF = 0;
constexpr int ST_DEBUG_SKIPB = 0;
constexpr int ST_DEBUG_SKIP = 0;
#endif

struct EpiParams {
  int64_t m_rows;  // valid rows of C
  int64_t n_cols;  // valid columns of C
  // --- STATS / DZ (LM head) ---
  const int64_t* ids;  // [T] global vocab ids
  int64_t T;           // tokens in the packed row; output index t = row + shift
  int32_t shift;
  int32_t round_bf16;
  int64_t col_base;  // global vocab id of column 0 of this GEMM (v_offset + chunk start)
  float inv_temp;
  float4* partials;        // STATS: [n_splits][part_stride] (m, s, u, z_target) per source row
  int64_t part_stride;
  const float* lse;        // DZ: [T]
  const float* dlogp;      // DZ: [T]
  __nv_bfloat16* dz;       // DZ: [rows][ldz] bf16 dlogits; STATS_STORE: [rows][ldz] bf16 logits
  int64_t ldz;
  // --- STORE ---
  float* c_f32;
  __nv_bfloat16* c_bf16;
  int64_t ldc;
  int32_t flags;
};

struct GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_c;  // output map for the TMA-store epilogues (fp32 {32 x 32} or bf16 {64 x 32} boxes, 128B swizzle)
  int32_t m_tiles, n_tiles, k_blocks;
  int32_t tiles_per_split, n_splits, group_m, num_items;
  int32_t max_stages;   // > 0: use at most this many of the smem ring's stages (prefetch-depth experiments)
  int32_t split_group;  // 0: m-tile fastest inside a band; S > 0: splits fastest in groups of S (neighbouring workers share an m-tile)
  int32_t k_splits;     // > 1: every output tile is produced by k_splits items over disjoint k ranges (epilogue must reduce-add)
  int32_t k_per_split;  // k-blocks per item (== k_blocks without k-splitting)
  uint64_t a_desc_hi, b_desc_hi;  // smem descriptor templates (everything but the start address)
  uint32_t a_kstep, b_kstep;      // byte advance of the start address per UMMA_K step
  uint64_t hint_a, hint_b, hint_c;  // L2 eviction priority of the A / B loads and of the epilogue's stores
  // Fused reduce-scatter of C over peer GPUs (vocab-sharded d_hidden): rows [j * peer_rows, (j + 1) * peer_rows)
  // of C belong to peer j and are reduce-added (fp32, TMA, performed at the owner's L2 over NVLink) into ITS accumulator through
  // tma_c_peer[j]; m_rot rotates the m-tile order so that the ranks do not all target the same owner at the same time.
  int32_t n_peers;
  int32_t peer_rows;
  int32_t m_rot;
  int32_t peer_row0;  // global row of row 0 of this GEMM's C (a vocab chunk of the d_weight GEMM starts inside some owner's range)
  int32_t peer_nrows[8];  // rows owner j holds (the last owners of a torch.chunk layout may hold fewer than peer_rows)
  float* peer_ptr[8];     // the same accumulators as raw pointers: slabs that straddle two owners leave through red.global
  CUtensorMap tma_c_peer[8];
  EpiParams epi;
};

struct Item {
  int m_tile, split, n_begin, n_count, k_begin, k_count;
};

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// (also compiled for the host: prl_debug_schedule lets the CPU tests check that a decomposition covers every tile once)
__host__ __device__ __forceinline__ Item decode_item(const GemmParams& p, int item) {
  const int ksp = p.k_splits > 1 ? p.k_splits : 1;  // k-splits of one tile are neighbouring items (they run concurrently)
  const int per_band = p.group_m * p.n_splits * ksp;
  const int band = item / per_band;
  const int m0 = band * p.group_m;
  const int rows = imin(p.group_m, p.m_tiles - m0);
  const int idx = item - band * per_band;
  Item it;
  if (p.split_group > 0) {
    // the workers that sweep the same vocab range are spread over the whole chip (both dies), neighbours share the A panel
    const int per_grp = rows * p.split_group;
    const int g = idx / per_grp;
    const int s_here = imin(p.split_group, p.n_splits - g * p.split_group);
    const int r = idx - g * per_grp;
    it.split = g * p.split_group + r % s_here;
    it.m_tile = m0 + r / s_here;
    it.k_begin = 0;
    it.k_count = p.k_blocks;
  } else {
    const int vs = idx / rows;  // (n split, k split)
    it.m_tile = m0 + idx % rows;
    if (p.m_rot > 0) it.m_tile = (it.m_tile + p.m_rot) % p.m_tiles;
    it.split = vs / ksp;
    it.k_begin = (vs - it.split * ksp) * p.k_per_split;
    it.k_count = imin(p.k_per_split, p.k_blocks - it.k_begin);
  }
  it.n_begin = it.split * p.tiles_per_split;
  it.n_count = imin(p.tiles_per_split, p.n_tiles - it.n_begin);
  return it;
}

// ------------------------------------------------------------------------------------------------
// epilogue pieces (one thread = one accumulator row)
// ------------------------------------------------------------------------------------------------
// Online softmax state of one accumulator row.  m, u and zt are kept in the ACCUMULATOR domain (a = bf16-rounded
// <hidden, W[v]>, before the 1/tau scaling) and converted when the partial is flushed: z = a / tau, so
// m_z = m / tau, u_z = u / tau, zt_z = zt / tau; s = sum 2^{(a - m) c1} with c1 = log2(e) / tau.
struct RowStats {
  float m, s, u, zt;
};

// bf16 rounding of a pair of accumulators, returned as fp32 (one F2FP + SHF + LOP3 for two elements)
__device__ __forceinline__ void round_pair(uint32_t v0, uint32_t v1, bool round, float& a0, float& a1) {
  if (round) {
    const uint32_t pk = ptx::cvt_bf16x2(__uint_as_float(v0), __uint_as_float(v1));
    a0 = __uint_as_float(pk << 16);
    a1 = __uint_as_float(pk & 0xFFFF0000u);
  } else {
    a0 = __uint_as_float(v0);
    a1 = __uint_as_float(v1);
  }
}

// 32 columns of one row: packed fp32x2 math (FFMA2 / FADD2), four independent accumulation chains, ~5 issue slots per logit
// STORE: also returns the 32 logits as 16 packed bf16x2 words (the rounding the statistics use anyway)
template <bool PARTIAL, bool STORE = false>
__device__ __forceinline__ void stats_chunk(const uint32_t (&v)[32], RowStats& st, float c1, bool round, int n_valid_in_chunk, int tj,
                                            uint32_t* w = nullptr) {
  float a[32];
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    if (STORE) {
      const uint32_t pk = ptx::cvt_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
      w[j >> 1] = pk;
      a[j] = __uint_as_float(pk << 16);
      a[j + 1] = __uint_as_float(pk & 0xFFFF0000u);
    } else {
      round_pair(v[j], v[j + 1], round, a[j], a[j + 1]);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    if (!PARTIAL || j < n_valid_in_chunk) mx0 = fmaxf(mx0, a[j]);
    if (!PARTIAL || j + 1 < n_valid_in_chunk) mx1 = fmaxf(mx1, a[j + 1]);
  }
  const float m_new = fmaxf(st.m, fmaxf(mx0, mx1));
  const float sc = ptx::ex2((st.m - m_new) * c1);  // st.m == -inf -> 0
  const float negm = -m_new * c1;
  const uint64_t C1 = ptx::pack2(c1, c1), NM = ptx::pack2(negm, negm);
  uint64_t S0 = ptx::pack2(st.s * sc, 0.f), S1 = 0, U0 = ptx::pack2(st.u * sc, 0.f), U1 = 0;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const uint64_t A0 = ptx::pack2(a[j], a[j + 1]), A1 = ptx::pack2(a[j + 2], a[j + 3]);
    float x0, x1, x2, x3;
    ptx::unpack2(ptx::fma2(A0, C1, NM), x0, x1);
    ptx::unpack2(ptx::fma2(A1, C1, NM), x2, x3);
    float e0 = ptx::ex2(x0), e1 = ptx::ex2(x1), e2 = ptx::ex2(x2), e3 = ptx::ex2(x3);
    if (PARTIAL) {
      e0 = (j < n_valid_in_chunk) ? e0 : 0.f;
      e1 = (j + 1 < n_valid_in_chunk) ? e1 : 0.f;
      e2 = (j + 2 < n_valid_in_chunk) ? e2 : 0.f;
      e3 = (j + 3 < n_valid_in_chunk) ? e3 : 0.f;
    }
    const uint64_t E0 = ptx::pack2(e0, e1), E1 = ptx::pack2(e2, e3);
    S0 = ptx::add2(S0, E0);
    S1 = ptx::add2(S1, E1);
    U0 = ptx::fma2(E0, A0, U0);
    U1 = ptx::fma2(E1, A1, U1);
  }
  float s0, s1, s2, s3, u0, u1, u2, u3;
  ptx::unpack2(S0, s0, s1);
  ptx::unpack2(S1, s2, s3);
  ptx::unpack2(U0, u0, u1);
  ptx::unpack2(U1, u2, u3);
  st.m = m_new;
  st.s = (s0 + s1) + (s2 + s3);
  st.u = (u0 + u1) + (u2 + u3);
  if ((unsigned)tj < 32u) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == tj) st.zt = a[j];
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// dz = coef * (1[j == tj] - exp(z - lse)) for 32 columns, as fp32 (c1 = log2(e) / tau, nlse2 = -lse * log2(e))
__device__ __forceinline__ void dz_values(const uint32_t (&v)[32], float c1, bool round, float coef, float nlse2, int tj, float (&d)[32]) {
  const uint64_t C1 = ptx::pack2(c1, c1), NL = ptx::pack2(nlse2, nlse2), NC = ptx::pack2(-coef, -coef);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float a0, a1, x0, x1;
    round_pair(v[j], v[j + 1], round, a0, a1);
    ptx::unpack2(ptx::fma2(ptx::pack2(a0, a1), C1, NL), x0, x1);
    ptx::unpack2(ptx::mul2(ptx::pack2(ptx::ex2(x0), ptx::ex2(x1)), NC), d[j], d[j + 1]);
  }
  if ((unsigned)tj < 32u) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == tj) d[j] += coef;
  }
}

// register-path store of one 32-column chunk of dlogits (PRL_EPI_TMA=0)
__device__ __forceinline__ void dz_chunk(const uint32_t (&v)[32], float c1, bool round, float coef, float nlse2, int tj,
                                         __nv_bfloat16* dst, int n_valid_in_chunk) {
  float d[32];
  dz_values(v, c1, round, coef, nlse2, tj, d);
  if (n_valid_in_chunk >= 32) {
    uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 w;
      w.x = ptx::cvt_bf16x2(d[8 * q + 0], d[8 * q + 1]);
      w.y = ptx::cvt_bf16x2(d[8 * q + 2], d[8 * q + 3]);
      w.z = ptx::cvt_bf16x2(d[8 * q + 4], d[8 * q + 5]);
      w.w = ptx::cvt_bf16x2(d[8 * q + 6], d[8 * q + 7]);
      o[q] = w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid_in_chunk) dst[j] = __float2bfloat16_rn(d[j]);
  }
}

// same math, result packed to 16 x bf16x2 for the TMA-store epilogue
__device__ __forceinline__ void dz_pack(const uint32_t (&v)[32], float c1, bool round, float coef, float nlse2, int tj, uint32_t (&w)[16]) {
  float d[32];
  dz_values(v, c1, round, coef, nlse2, tj, d);
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = ptx::cvt_bf16x2(d[2 * j], d[2 * j + 1]);
}

// register-path store of 32 bf16 values held as 16 packed words (PRL_EPI_TMA=0)
__device__ __forceinline__ void store_words(const uint32_t (&w)[16], __nv_bfloat16* dst, int n_valid_in_chunk) {
  if (n_valid_in_chunk >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_uint4(w[4 * q + 0], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(dst);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid_in_chunk) o[j] = (uint16_t)((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu));
  }
}

__device__ __forceinline__ void store_chunk(const uint32_t (&v)[32], const EpiParams& e, int64_t row, int64_t col0, int n_valid_in_chunk) {
  float c[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) c[j] = __uint_as_float(v[j]);
  const int64_t off = row * e.ldc + col0;
  const bool vec = (n_valid_in_chunk >= 32) && ((e.ldc & 7) == 0);
  if (vec) {
    if (e.flags & ST_LOAD_F32) {
      const float4* src = reinterpret_cast<const float4*>(e.c_f32 + off);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = src[q];
        c[4 * q + 0] += x.x;
        c[4 * q + 1] += x.y;
        c[4 * q + 2] += x.z;
        c[4 * q + 3] += x.w;
      }
    }
    if (e.flags & ST_LOAD_BF16) {
      const uint4* src = reinterpret_cast<const uint4*>(e.c_bf16 + off);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 x = src[q];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          c[8 * q + 2 * k + 0] += __uint_as_float(w[k] << 16);
          c[8 * q + 2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
        }
      }
    }
    if (e.flags & ST_STORE_F32) {
      float4* dst = reinterpret_cast<float4*>(e.c_f32 + off);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = make_float4(c[4 * q + 0], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
    }
    if (e.flags & ST_STORE_BF16) {
      uint4* dst = reinterpret_cast<uint4*>(e.c_bf16 + off);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(c[8 * q + 0], c[8 * q + 1]);
        w.y = pack_bf16x2(c[8 * q + 2], c[8 * q + 3]);
        w.z = pack_bf16x2(c[8 * q + 4], c[8 * q + 5]);
        w.w = pack_bf16x2(c[8 * q + 6], c[8 * q + 7]);
        dst[q] = w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < n_valid_in_chunk) {
        float x = c[j];
        if (e.flags & ST_LOAD_F32) x += e.c_f32[off + j];
        if (e.flags & ST_LOAD_BF16) x += __bfloat162float(e.c_bf16[off + j]);
        if (e.flags & ST_STORE_F32) e.c_f32[off + j] = x;
        if (e.flags & ST_STORE_BF16) e.c_bf16[off + j] = __float2bfloat16_rn(x);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel.  CG = 1: one CTA per 128 x 256 tile (cta_group::1).  CG = 2: a CTA pair (cluster of 2, cta_group::2)
// per 256 x 256 tile: each CTA stages its own 128 rows of A and its own half (128 rows) of B, the leader CTA
// issues M = 256 MMAs that read both CTAs' shared memory, each CTA's TMEM holds its 128 accumulator rows.
// Halving the B bytes per SM and k-step gives 6 instead of 4 pipeline stages in the same shared memory.
// CG = 4: a cluster of two such pairs stacked along M (512 x 256 per cluster) that share the B tile: every CTA
// fetches only a quarter of B (64 rows) from L2 and TMA-multicasts it to its counterpart in the other pair, so the
// L2 -> SM operand traffic per MMA drops by another 25 % (the path is power-bound; section 7 of DESIGN.md).
// ------------------------------------------------------------------------------------------------
// MS = 2 (CG = 2 only): TWO 256-row M sub-tiles per CTA pair share every B k-block (512 x 256 per pair).  A's bytes per
// MMA stay the same, B's halve: 25 % fewer operand bytes from L2 per MMA, which matters because the path is power-bound
// (DESIGN.md section 7).  The two sub-tiles use the two 256-column TMEM accumulators, so the epilogue (8 warps: one group
// of 4 per sub-tile) no longer overlaps the next tile's MMAs; the trade is measured in DESIGN.md.
template <int CG, int MS, int EPI>
struct Cfg {
  static constexpr int kPairs = (CG == 4) ? 2 : 1;               // CTA pairs per cluster
  static constexpr int kBRows = BLOCK_N / (CG >= 2 ? 2 : 1);     // B rows staged in one CTA's smem
  static constexpr int kBLoad = kBRows / kPairs;                 // B rows one CTA fetches itself (the rest arrives by multicast)
  static constexpr int kASub = BLOCK_M * BLOCK_K * 2;            // 16 KB per M sub-tile
  static constexpr int kAStage = MS * kASub;
  static constexpr int kBStage = kBRows * BLOCK_K * 2;           // 32 KB / 16 KB
  static constexpr int kStage = kAStage + kBStage;
  static constexpr int kEpiWarps = 4 * MS;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  // smem staging of the TMA-store epilogues: 2 buffers of 32 rows x 128 B per epilogue warp (none for the statistics epilogue)
  static constexpr int kStaging = (EPI == EPI_STATS) ? 0 : kEpiWarps * 2 * 4096;
  static constexpr int kBudget = 232448 - 1024 /*align slack*/ - 256 /*barriers*/ - kStaging;
  static constexpr int kFit = kBudget / kStage;
  static constexpr int kStages = (CG == 1) ? PRL_STAGES : (kFit > 8 ? 8 : kFit);
  static constexpr int kSmem = kStages * kStage + kStaging + 1024 + 256;
  static constexpr int kRowsPerTile = BLOCK_M * CG * MS;         // rows of C one cluster produces per tile
};

template <int CG, int MS, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(Cfg<CG, MS, EPI>::kThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<CG, MS, EPI>;
  constexpr int NS = C::kStages;
  const int ns_run = (p.max_stages > 0 && p.max_stages < NS) ? p.max_stages : NS;  // stages actually cycled through
  static_assert(MS == 1 || CG == 2, "M sub-tiling is implemented for CTA pairs only");
  static_assert(NS >= 2, "pipeline needs at least two stages");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (the dynamic smem base offset is the same in both CTAs of a pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NS * C::kAStage;
  uint8_t* smem_stage = smem + NS * C::kStage;  // 1024-byte aligned: the manual 128B swizzle below relies on it
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * C::kStage + C::kStaging);
  uint64_t* full_bar = bars;                  // [NS]   TMA -> MMA          (CG >= 2: the pair leader's is used)
  uint64_t* empty_bar = bars + NS;            // [NS]   MMA -> TMA          (per CTA)
  uint64_t* tfull_bar = bars + 2 * NS;        // [2]    MMA -> epilogue     (per CTA)
  uint64_t* tempty_bar = bars + 2 * NS + 2;   // [2]    epilogue -> MMA     (CG >= 2: the pair leader's is used)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool kPair = CG >= 2;                                   // cta_group::2 MMAs
  const int crank = kPair ? (int)ptx::cluster_ctarank() : 0;        // rank in the cluster: 0 .. CG-1
  const int rank = crank & 1;                                       // rank in the CTA pair (0 = leader)
  const int pair = crank >> 1;                                      // which pair of the cluster
  const int worker = blockIdx.x / CG;                               // one worker = one cluster
  const int n_workers = gridDim.x / CG;
  constexpr int kSubRows = BLOCK_M * (kPair ? 2 : 1) * C::kPairs;   // rows between consecutive M sub-tiles of a tile

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tma_a);
    ptx::prefetch_tensormap(&p.tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], C::kPairs);  // every pair that reads data this CTA wrote must release the slot
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], kPair ? 8 : 4);  // one arrive per epilogue warp of the group (of both CTAs of the pair)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) {
      ptx::tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
      ptx::tmem_relinquish_cg2();
    } else {
      ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kPair) ptx::cluster_sync_all();  // the peers' barriers must be initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (every CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < p.num_items; item += n_workers) {
        const Item it = decode_item(p, item);
        const int m0 = it.m_tile * C::kRowsPerTile + crank * BLOCK_M;
        for (int nt = it.n_begin; nt < it.n_begin + it.n_count; ++nt) {
          const int n0 = nt * BLOCK_N + rank * C::kBRows;
          for (int kb = it.k_begin; kb < it.k_begin + it.k_count; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            // profiling aid (PRL_DEBUG_SKIPB=1): skip the B load on odd k-blocks (stale data, WRONG results) to measure how
            // much of the power budget the L2->SM operand traffic costs
            const bool skip_b = (p.epi.flags & ST_DEBUG_SKIPB) && (kb & 1);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], (skip_b ? C::kAStage : C::kStage) * (kPair ? 2 : 1));
            uint8_t* sa = smem_a + stage * C::kAStage;
            uint8_t* sb = smem_b + stage * C::kBStage;
            auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1, uint64_t hint) {
              if (kPair) {
                if (hint == 0)
                  ptx::tma_load_2d_cg2(dst, map, &full_bar[stage], c0, c1);
                else
                  ptx::tma_load_2d_cg2_hint(dst, map, &full_bar[stage], c0, c1, hint);
              } else {
                if (hint == 0)
                  ptx::tma_load_2d(dst, map, &full_bar[stage], c0, c1);
                else
                  ptx::tma_load_2d_hint(dst, map, &full_bar[stage], c0, c1, hint);
              }
            };
            // B: with two pairs per cluster this CTA fetches quarter `pair` of the tile and multicasts it to the CTA of
            // the same pair-rank in both pairs (cluster ranks rank and rank + 2)
            auto load_b = [&](void* dst, int c0, int c1) {
              if (CG == 4)
                ptx::tma_load_2d_cg2_mc_hint(dst, &p.tma_b, &full_bar[stage], c0, c1, (uint16_t)((1u << rank) | (4u << rank)),
                                             p.hint_b ? p.hint_b : ptx::kEvictNormal);
              else
                load(dst, &p.tma_b, c0, c1, p.hint_b);
            };
#pragma unroll
            for (int ms = 0; ms < MS; ++ms) {
              uint8_t* sas = sa + ms * C::kASub;
              const int m0s = m0 + ms * kSubRows;
              if (!A_MN) {
                load(sas, &p.tma_a, kb * BLOCK_K, m0s, p.hint_a);
              } else {
#pragma unroll
                for (int s = 0; s < BLOCK_M / 64; ++s) load(sas + s * (BLOCK_K * 128), &p.tma_a, m0s + s * 64, kb * BLOCK_K, p.hint_a);
              }
            }
            if (skip_b) {
            } else if (!B_MN) {
              load_b(sb + pair * (C::kBLoad * 128), kb * BLOCK_K, n0 + pair * C::kBLoad);
            } else {
#pragma unroll
              for (int s = 0; s < C::kBLoad / 64; ++s) {
                const int slab = pair * (C::kBLoad / 64) + s;
                load_b(sb + slab * (BLOCK_K * 128), n0 + slab * 64, kb * BLOCK_K);
              }
            }
            if (++stage == ns_run) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M * (kPair ? 2 : 1), BLOCK_N, A_MN, B_MN);
      const uint16_t mask_all = (uint16_t)((1u << CG) - 1u), mask_pair = (uint16_t)(3u << (2 * pair));
      // smem descriptor templates and k-step advances come from the host (make_desc_defaults)
      const uint64_t a_hi = p.a_desc_hi, b_hi = p.b_desc_hi;
      const uint32_t a_kstep = p.a_kstep, b_kstep = p.b_kstep;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;             // MS == 1: accumulator buffer of the current tile (double buffering)
      uint32_t acc_phase = 0;  // MS == 2: both buffers belong to every tile, the phase flips per tile
      for (int item = worker; item < p.num_items; item += n_workers) {
        const Item it = decode_item(p, item);
        for (int nt = 0; nt < it.n_count; ++nt) {
          if (MS == 1) {
            ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          } else {
            ptx::mbar_wait(&tempty_bar[0], acc_phase ^ 1);
            ptx::mbar_wait(&tempty_bar[1], acc_phase ^ 1);
          }
          ptx::tc_fence_after();
          for (int kb = 0; kb < it.k_count; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kAStage);
            const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBStage);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t bd = ptx::smem_desc(b_hi, sb + k * b_kstep);
#pragma unroll
              for (int ms = 0; ms < MS; ++ms) {
                const uint64_t ad = ptx::smem_desc(a_hi, sa + ms * C::kASub + k * a_kstep);
                const uint32_t d_tmem = tmem_base + (MS == 1 ? acc : ms) * BLOCK_N;
                if (kPair)
                  ptx::umma_bf16_cg2(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                else
                  ptx::umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            // frees the smem slot when these MMAs retire: in every CTA of the cluster that wrote part of it
            if (kPair)
              ptx::umma_commit_cg2(&empty_bar[stage], mask_all);
            else
              ptx::umma_commit(&empty_bar[stage]);
            if (++stage == ns_run) {
              stage = 0;
              phase ^= 1;
            }
          }
          // accumulator(s) ready for the epilogue warps (of both CTAs of this pair)
#pragma unroll
          for (int ms = 0; ms < MS; ++ms) {
            uint64_t* bar = &tfull_bar[MS == 1 ? acc : ms];
            if (kPair)
              ptx::umma_commit_cg2(bar, mask_pair);
            else
              ptx::umma_commit(bar);
          }
          if (MS == 1) {
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          } else {
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (every CTA: its own 128 accumulator rows per sub-tile) =====================
    const int quarter = warp & 3;
    const int group = (warp - 4) >> 2;  // M sub-tile this warp group drains (MS == 2); 0 otherwise
    const int ewarp = warp - 4;         // index of this warp's staging buffers
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const EpiParams& e = p.epi;
    const float c1 = e.inv_temp * LOG2E;  // accumulator -> log2-domain logit
    int acc = (MS == 1) ? 0 : group;
    uint32_t acc_phase = 0;
    int stage_buf = 0;  // which of this warp's two staging buffers is filled next
    for (int item = worker; item < p.num_items; item += n_workers) {
      const Item it = decode_item(p, item);
      const int64_t row_base = (int64_t)it.m_tile * C::kRowsPerTile + group * kSubRows + crank * BLOCK_M;
      const int64_t row = row_base + row_in_tile;
      const bool row_ok = row < e.m_rows;
      // per-item row state
      RowStats st = {-INFINITY, 0.f, 0.f, 0.f};
      int64_t tcol = -1;  // target column of this row, local to this GEMM's column space
      // rows without a target (t >= T) keep coef = 0 and exp2(.. - inf) = 0, so their dlogits are exactly 0
      float coef = 0.f, nlse2 = -INFINITY;
      if (EPI == EPI_STATS || EPI == EPI_STATS_STORE || EPI == EPI_DZ) {
        const int64_t t = row + e.shift;
        if (row_ok && t < e.T) {
          tcol = e.ids[t] - e.col_base;
          if (EPI == EPI_DZ) {
            coef = e.dlogp[t] * e.inv_temp;
            nlse2 = -e.lse[t] * LOG2E;
          }
        }
      }
      for (int nt = it.n_begin; nt < it.n_begin + it.n_count; ++nt) {
        if (p.epi.flags & ST_WAIT_BACKOFF)
          ptx::mbar_wait_backoff(&tfull_bar[acc], acc_phase, 256);
        else
          ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
        const int64_t n0 = (int64_t)nt * BLOCK_N;
        const int64_t n_left = e.n_cols - n0;  // > 0
        const bool use_tma = (EPI != EPI_STATS) && (e.flags & ST_TMA) && !(e.flags & ST_DEBUG_SKIP);
        const bool out_bf16 = (EPI == EPI_DZ) || (EPI == EPI_STATS_STORE) || (e.flags & ST_STORE_BF16);
        const int row0_warp = (int)row_base + quarter * 32;
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
          const int64_t c0 = n0 + ch * 32;
          const int64_t valid64 = n_left - ch * 32;
          if (valid64 <= 0) break;  // uniform across the CTA
          const int n_valid = valid64 >= 32 ? 32 : (int)valid64;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + acc * BLOCK_N + ch * 32, v);
          ptx::tmem_ld_wait();
          const int64_t tj64 = tcol - c0;
          const int tj = (tj64 >= 0 && tj64 < 32) ? (int)tj64 : -1;
          if (EPI == EPI_STORE && (e.flags & ST_ADD_LOCAL) && row_ok) {
            const float* src = e.c_f32 + row * e.ldc + c0;
            if (n_valid == 32) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 x = __ldcs(reinterpret_cast<const float4*>(src) + q);  // read once, streaming
                v[4 * q + 0] = __float_as_uint(__uint_as_float(v[4 * q + 0]) + x.x);
                v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + x.y);
                v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + x.z);
                v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + x.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < n_valid) v[j] = __float_as_uint(__uint_as_float(v[j]) + src[j]);
            }
          }
          if (EPI == EPI_STATS) {
            if (e.flags & ST_DEBUG_SKIP) {
              st.s += __uint_as_float(v[0]) + __uint_as_float(v[31]);  // profiling aid: keep the TMEM load alive, skip the statistics
            } else if (n_valid == 32)
              stats_chunk<false>(v, st, c1, e.round_bf16 != 0, 32, tj);
            else
              stats_chunk<true>(v, st, c1, e.round_bf16 != 0, n_valid, tj);
          } else if (EPI == EPI_STATS_STORE && !use_tma) {
            uint32_t w[16];
            if (n_valid == 32)
              stats_chunk<false, true>(v, st, c1, true, 32, tj, w);
            else
              stats_chunk<true, true>(v, st, c1, true, n_valid, tj, w);
            if (row_ok) store_words(w, e.dz + row * e.ldz + c0, n_valid);
          } else if (!use_tma) {
            if (EPI == EPI_DZ) {
              if (row < e.T && !(e.flags & ST_DEBUG_SKIP))
                dz_chunk(v, c1, e.round_bf16 != 0, coef, nlse2, tj, e.dz + row * e.ldz + c0, n_valid);
            } else {
              if (row_ok && !(e.flags & ST_DEBUG_SKIP)) store_chunk(v, e, row, c0, n_valid);
            }
          } else if (EPI == EPI_STORE && p.n_peers > 0 && (p.peer_row0 + row0_warp) % p.peer_rows + 32 > p.peer_rows &&
                     (p.peer_row0 + row0_warp) / p.peer_rows + 1 < p.n_peers) {
            // ---- fused reduce-scatter, a 32-row slab that straddles two owners (peer_rows not a multiple of 32; at most one slab
            // per owner boundary and column chunk): every lane adds its own row into whichever owner holds it, 16 bytes per
            // red.global (a TMA box cannot be split, and negative box coordinates are not accepted by the reduce form)
            const int g = p.peer_row0 + row0_warp + lane;
            const int own = g / p.peer_rows;
            const int loc = g - own * p.peer_rows;
            if (own < p.n_peers && loc < p.peer_nrows[own]) {
              float* dst = p.peer_ptr[own] + (int64_t)loc * e.ldc + c0;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (4 * q + 3 < n_valid) {
                  ptx::red_add_v4_f32(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                      __uint_as_float(v[4 * q + 3]));
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (4 * q + k < n_valid) atomicAdd(dst + 4 * q + k, __uint_as_float(v[4 * q + k]));
                }
              }
            }
          } else if (C::kStaging > 0) {
            // ---- registers -> 128B-swizzled smem staging -> TMA store / reduce-add (full-line writes, no SM-side reads)
            // fp32: one 32 x 32 box per chunk.  bf16: two chunks share one 64 x 32 box (128 B per row).
            const bool second_half = out_bf16 && (ch & 1);
            if (!second_half) {
              // the buffer about to be refilled was handed to TMA two stores ago: wait until it has been read
              if (lane == 0) ptx::bulk_wait_read<1>();
              __syncwarp();
            }
            const uint32_t sbuf = ptx::smem_u32(smem_stage) + (uint32_t)((ewarp * 2 + stage_buf) * 4096) + (uint32_t)lane * 128u;
            const uint32_t sw = (uint32_t)(lane & 7);
            if (!out_bf16) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                ptx::st_shared_v4(sbuf + (((uint32_t)q ^ sw) << 4), v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              uint32_t w[16];
              if (EPI == EPI_DZ) {
                dz_pack(v, c1, e.round_bf16 != 0, coef, nlse2, tj, w);
              } else if (EPI == EPI_STATS_STORE) {
                if (n_valid == 32)
                  stats_chunk<false, true>(v, st, c1, true, 32, tj, w);
                else
                  stats_chunk<true, true>(v, st, c1, true, n_valid, tj, w);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
              }
              const uint32_t half = (uint32_t)(ch & 1) * 4u;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                ptx::st_shared_v4(sbuf + (((half + (uint32_t)q) ^ sw) << 4), w[4 * q + 0], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
            const bool last_chunk = (valid64 <= 32) || (ch == BLOCK_N / 32 - 1);
            if (!out_bf16 || second_half || last_chunk) {
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                const void* src = smem_stage + (ewarp * 2 + stage_buf) * 4096;
                const int32_t col = (int32_t)(out_bf16 ? (n0 + (ch & ~1) * 32) : c0);
                const bool reduce = (EPI == EPI_STORE) && (e.flags & ST_LOAD_F32);
                if (EPI == EPI_STORE && p.n_peers > 0) {
                  // fused reduce-scatter: add this 32-row slab into its owner's fp32 accumulator (local or NVLink peer memory);
                  // rows beyond the owner's range are clipped by its tensor map (slabs that straddle two owners took the
                  // red.global path above)
                  const int g0 = p.peer_row0 + row0_warp;
                  const int owner = g0 / p.peer_rows;
                  if (owner < p.n_peers) ptx::tma_reduce_add_2d(&p.tma_c_peer[owner], src, col, g0 - owner * p.peer_rows);
                } else if (p.hint_c == 0) {
                  if (reduce)
                    ptx::tma_reduce_add_2d(&p.tma_c, src, col, row0_warp);
                  else
                    ptx::tma_store_2d(&p.tma_c, src, col, row0_warp);
                } else {
                  if (reduce)
                    ptx::tma_reduce_add_2d_hint(&p.tma_c, src, col, row0_warp, p.hint_c);
                  else
                    ptx::tma_store_2d_hint(&p.tma_c, src, col, row0_warp, p.hint_c);
                }
                ptx::bulk_commit();
              }
              stage_buf ^= 1;
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair)
            ptx::mbar_arrive_leader(&tempty_bar[acc]);
          else
            ptx::mbar_arrive(&tempty_bar[acc]);
        }
        if (MS == 1) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
      if (EPI == EPI_STATS || EPI == EPI_STATS_STORE) {
        if (row_ok)
          e.partials[(int64_t)it.split * e.part_stride + row] = make_float4(st.m * e.inv_temp, st.s, st.u * e.inv_temp, st.zt * e.inv_temp);
      }
    }
    if (EPI != EPI_STATS && lane == 0) ptx::bulk_wait_all();  // staging smem must outlive the last TMA store
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (kPair) ptx::cluster_sync_all();  // no CTA of the cluster may exit while another can still signal it
  if (warp == 2) {
    ptx::tc_fence_after();
    if (kPair)
      ptx::tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else
      ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace prl
