// HBM/latency-bound kernels of the RL loss path: group advantages, segment boundaries, packed clipped
// importance-ratio loss (+ its gradient), and the API-compat row statistics on materialised logits.
// Reference behaviour (file:line relative to the reference root) is cited at each kernel.
#include <math.h>

#include "common.cuh"

namespace prl {

ErrState& err_state() {
  static thread_local ErrState s = {{0}, 0};
  return s;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

constexpr unsigned FULL = 0xffffffffu;

// torch.clamp(max=) / torch.maximum / torch.min / torch.max PROPAGATE NaN, fminf / fmaxf drop it: these keep the reference's
// behaviour (a NaN log-prob must surface as a NaN loss, not as a silently clipped ratio)
__device__ __forceinline__ float clamp_max_nan(float x, float hi) { return (x != x) ? x : fminf(x, hi); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }
__device__ __forceinline__ double min_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmin(a, b)); }
__device__ __forceinline__ double max_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmax(a, b)); }

// ------------------------------------------------------------------------------------------------
// advantages: one warp per group.  src/prime_rl/orchestrator/advantage.py:18,32,41-43,55-57.
// fp32 sum order ("L8"): lane accumulators acc[j] = x[j] + x[j+8] + ... (j < 8, left to right), then
// acc[0] + acc[1] + ... + acc[7], then the len%8 tail left to right.  Explicit __fadd_rn/__fmul_rn/
// __fdiv_rn keep nvcc from contracting into FMAs, so the result is bit-identical to the oracle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_l8(const float* __restrict__ x, int G, int lane, bool weighted,
                                             const int64_t* __restrict__ w, bool squared_dev = false, float center = 0.f) {
  const int n8 = (G / 8) * 8;
  float acc = 0.f;
  bool first = true;
  for (int c = 0; c < n8; c += 32) {
    const int i = c + lane;
    float v = 0.f;
    if (i < n8) {
      v = x[i];
      if (weighted) v = __fmul_rn(v, (float)w[i]);
      if (squared_dev) {
        const float d = __fsub_rn(v, center);
        v = __fmul_rn(d, d);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = __shfl_sync(FULL, v, (lane & 7) + 8 * q);
      if (c + 8 * q < n8) {
        acc = first ? t : __fadd_rn(acc, t);
        first = false;
      }
    }
  }
  float s;
  if (n8 > 0) {
    s = __shfl_sync(FULL, acc, 0);
#pragma unroll
    for (int j = 1; j < 8; ++j) s = __fadd_rn(s, __shfl_sync(FULL, acc, j));
  } else {
    s = 0.f;
  }
  for (int i = n8; i < G; ++i) {
    float v = x[i];
    if (weighted) v = __fmul_rn(v, (float)w[i]);
    if (squared_dev) {
      const float d = __fsub_rn(v, center);
      v = __fmul_rn(d, d);
    }
    s = __fadd_rn(s, v);
  }
  return s;
}

// GRPO's std-normalised advantage (NOT in the reference's registry, which is Dr.GRPO = mean only; north_star names "mean/std"):
//   a_i = (r_i - mean) / (std + PRL_ADV_GRPO_EPS),  std = sqrt(sum (r_i - mean)^2 / (G - 1))   (unbiased, as torch.std)
// with both sums in the same L8 order as the mean.  G == 1 gives nan, as torch.std of one element does.
__device__ __forceinline__ float grpo_scale(float ssq, int G) {
  const float std = __fsqrt_rn(__fdiv_rn(ssq, (float)(G - 1)));
  return __fadd_rn(std, PRL_ADV_GRPO_EPS);
}

// G <= 32: floor(32 / G) groups per warp, one rollout per lane, fully coalesced loads and stores.  Every lane of a
// group evaluates the same L8-ordered sum with ~G/8 + 8 shuffles inside its sub-warp (redundant but branch-free): 8 B of HBM
// traffic per rollout (16 B for opo), so large batches run at HBM speed instead of one under-filled warp per group.
constexpr int ADV_U = 4;  // independent group-sets in flight per warp (memory- and shuffle-latency hiding)

__device__ __forceinline__ void subwarp_sum_l8(const float (&x)[ADV_U], int G, int base, float (&s)[ADV_U]) {
  const int n8 = (G / 8) * 8;
  // lane accumulators: lane j < 8 of the group gets ((x[j] + x[j+8]) + x[j+16]) + x[j+24]  (other lanes: unused values)
  float a[ADV_U];
#pragma unroll
  for (int u = 0; u < ADV_U; ++u) a[u] = x[u];
  for (int q = 8; q < n8; q += 8) {
#pragma unroll
    for (int u = 0; u < ADV_U; ++u) a[u] = __fadd_rn(a[u], __shfl_down_sync(FULL, x[u], q));
  }
#pragma unroll
  for (int u = 0; u < ADV_U; ++u) s[u] = 0.f;
  if (n8 > 0) {
#pragma unroll
    for (int u = 0; u < ADV_U; ++u) s[u] = __shfl_sync(FULL, a[u], base);
#pragma unroll
    for (int j = 1; j < 8; ++j) {
#pragma unroll
      for (int u = 0; u < ADV_U; ++u) s[u] = __fadd_rn(s[u], __shfl_sync(FULL, a[u], base + j));
    }
  }
  for (int i = n8; i < G; ++i) {
#pragma unroll
    for (int u = 0; u < ADV_U; ++u) s[u] = __fadd_rn(s[u], __shfl_sync(FULL, x[u], base + i));
  }
}

__global__ void __launch_bounds__(256) advantages_small_group_kernel(const float* __restrict__ rewards,
                                                                     const int64_t* __restrict__ lengths, int64_t ngroups, int G,
                                                                     int kind, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gpw = 32 / G;                      // groups per warp and group-set
  const int gl = lane / G;                     // group of this lane inside the warp (>= gpw: idle lane)
  const int base = gl < gpw ? gl * G : 0;      // idle lanes still take part in the shuffles
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_sets = (ngroups + gpw - 1) / gpw;
  const int64_t n_iter = (n_sets + ADV_U - 1) / ADV_U;
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < n_iter; it += warps_total) {
    int64_t idx[ADV_U];
    bool active[ADV_U];
    float r[ADV_U], w[ADV_U], num[ADV_U], den[ADV_U];
#pragma unroll
    for (int u = 0; u < ADV_U; ++u) {
      const int64_t g = (it * ADV_U + u) * gpw + gl;
      active[u] = gl < gpw && g < ngroups;
      idx[u] = g * G + (lane - gl * G);
      r[u] = active[u] ? rewards[idx[u]] : 0.f;
    }
    if (kind == PRL_ADV_OPO) {
      float rw[ADV_U];
#pragma unroll
      for (int u = 0; u < ADV_U; ++u) {
        w[u] = active[u] ? (float)lengths[idx[u]] : 0.f;
        rw[u] = __fmul_rn(r[u], w[u]);
      }
      subwarp_sum_l8(rw, G, base, num);
      subwarp_sum_l8(w, G, base, den);
    } else {
      subwarp_sum_l8(r, G, base, num);
#pragma unroll
      for (int u = 0; u < ADV_U; ++u) den[u] = (float)G;
    }
    float scale[ADV_U];
    if (kind == PRL_ADV_GRPO) {
      float sq[ADV_U], ssq[ADV_U];
#pragma unroll
      for (int u = 0; u < ADV_U; ++u) {
        const float d = __fsub_rn(r[u], __fdiv_rn(num[u], den[u]));
        sq[u] = __fmul_rn(d, d);
      }
      subwarp_sum_l8(sq, G, base, ssq);
#pragma unroll
      for (int u = 0; u < ADV_U; ++u) scale[u] = grpo_scale(ssq[u], G);
    }
#pragma unroll
    for (int u = 0; u < ADV_U; ++u) {
      if (active[u]) {
        float a = __fsub_rn(r[u], __fdiv_rn(num[u], den[u]));
        if (kind == PRL_ADV_DRGRPO_NEGCLIPPED) a = max_nan(a, 0.f);
        if (kind == PRL_ADV_RLOO) a = __fdiv_rn(__fmul_rn(a, (float)G), (float)(G - 1));
        if (kind == PRL_ADV_GRPO) a = __fdiv_rn(a, scale[u]);
        out[idx[u]] = a;
      }
    }
  }
}

// G in {8, 16, 24, 32} (every BASELINE configuration): each lane owns EIGHT consecutive rollouts (two 16-byte loads, two
// 16-byte stores), M = G / 8 consecutive lanes own a group.  The L8 lane accumulator j of the group is
// ((x[j] + x[j+8]) + x[j+16]) + x[j+24] = the j-th component of lanes 0..M-1 added in lane order (M shuffles per component), the
// total is the eight accumulators added left to right: the same fp32 operations in the same order as the generic kernels
// (and the oracle), ~5 instead of ~20 instructions per rollout, so large batches run at HBM speed.
template <int M>
__global__ void __launch_bounds__(256) advantages_vec8_kernel(const float* __restrict__ rewards, const int64_t* __restrict__ lengths,
                                                              int64_t ngroups, int kind, float* __restrict__ out) {
  constexpr int G = 8 * M;
  constexpr int GPW = 32 / M;  // groups per warp (M = 3: 10 groups, 2 idle lanes)
  const int lane = threadIdx.x & 31;
  const int gl = lane / M, q = lane - gl * M;
  const int base = gl < GPW ? gl * M : 0;  // idle lanes still take part in the shuffles
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_sets = (ngroups + GPW - 1) / GPW;
  for (int64_t set = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); set < n_sets; set += warps_total) {
    const int64_t g = set * GPW + gl;
    const bool active = gl < GPW && g < ngroups;
    const int64_t e0 = g * G + 8 * q;  // first of this lane's eight rollouts
    float x[8], w[8];
    if (active) {
      const float4 a = *reinterpret_cast<const float4*>(rewards + e0), b = *reinterpret_cast<const float4*>(rewards + e0 + 4);
      x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
    }
    float num, den = (float)G;
    auto sum_l8 = [&](const float (&v)[8]) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = __shfl_sync(FULL, v[j], base);
#pragma unroll
        for (int m = 1; m < M; ++m) acc[j] = __fadd_rn(acc[j], __shfl_sync(FULL, v[j], base + m));
      }
      float t = acc[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) t = __fadd_rn(t, acc[j]);
      return t;
    };
    if (kind == PRL_ADV_OPO) {
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const longlong2 l2 = *reinterpret_cast<const longlong2*>(lengths + e0 + j);
          w[j] = (float)l2.x, w[j + 1] = (float)l2.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 0.f;
      }
      float rw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) rw[j] = __fmul_rn(x[j], w[j]);
      num = sum_l8(rw);
      den = sum_l8(w);
    } else {
      num = sum_l8(x);
    }
    const float mean = __fdiv_rn(num, den);
    float scale = 1.f;
    if (kind == PRL_ADV_GRPO) {  // every lane takes part in the shuffles of the second sum
      float sq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = __fsub_rn(x[j], mean);
        sq[j] = __fmul_rn(d, d);
      }
      scale = grpo_scale(sum_l8(sq), G);
    }
    if (active) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = __fsub_rn(x[j], mean);
        if (kind == PRL_ADV_DRGRPO_NEGCLIPPED) a = max_nan(a, 0.f);
        if (kind == PRL_ADV_RLOO) a = __fdiv_rn(__fmul_rn(a, (float)G), (float)(G - 1));
        if (kind == PRL_ADV_GRPO) a = __fdiv_rn(a, scale);
        o[j] = a;
      }
      *reinterpret_cast<float4*>(out + e0) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(out + e0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

__global__ void __launch_bounds__(256) advantages_kernel(const float* __restrict__ rewards, const int64_t* __restrict__ lengths,
                                                         int64_t ngroups, int G, int kind, float* __restrict__ out,
                                                         const int64_t* __restrict__ tok_offsets, float* __restrict__ token_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ngroups; g += warps_total) {
    const float* r = rewards + g * G;
    float base;
    if (kind == PRL_ADV_OPO) {
      const int64_t* w = lengths + g * G;
      const float num = warp_sum_l8(r, G, lane, true, w);
      // weights.sum(): same order on the fp32-converted lengths
      const int n8 = (G / 8) * 8;
      float acc = 0.f;
      bool first = true;
      for (int c = 0; c < n8; c += 32) {
        const int i = c + lane;
        float v = (i < n8) ? (float)w[i] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float t = __shfl_sync(FULL, v, (lane & 7) + 8 * q);
          if (c + 8 * q < n8) {
            acc = first ? t : __fadd_rn(acc, t);
            first = false;
          }
        }
      }
      float den = 0.f;
      if (n8 > 0) {
        den = __shfl_sync(FULL, acc, 0);
#pragma unroll
        for (int j = 1; j < 8; ++j) den = __fadd_rn(den, __shfl_sync(FULL, acc, j));
      }
      for (int i = n8; i < G; ++i) den = __fadd_rn(den, (float)w[i]);
      base = __fdiv_rn(num, den);
    } else {
      base = __fdiv_rn(warp_sum_l8(r, G, lane, false, nullptr), (float)G);
    }
    float scale = 1.f;
    if (kind == PRL_ADV_GRPO) scale = grpo_scale(warp_sum_l8(r, G, lane, false, nullptr, true, base), G);
    for (int i = lane; i < G; i += 32) {
      float a = __fsub_rn(r[i], base);
      if (kind == PRL_ADV_DRGRPO_NEGCLIPPED) a = max_nan(a, 0.f);
      if (kind == PRL_ADV_RLOO) a = __fdiv_rn(__fmul_rn(a, (float)G), (float)(G - 1));
      if (kind == PRL_ADV_GRPO) a = __fdiv_rn(a, scale);
      out[g * G + i] = a;
    }
    if (token_out != nullptr) {
      __syncwarp();
      for (int i = 0; i < G; ++i) {
        const int64_t idx = g * G + i;
        float a = __fsub_rn(r[i], base);
        if (kind == PRL_ADV_DRGRPO_NEGCLIPPED) a = max_nan(a, 0.f);
        if (kind == PRL_ADV_RLOO) a = __fdiv_rn(__fmul_rn(a, (float)G), (float)(G - 1));
        if (kind == PRL_ADV_GRPO) a = __fdiv_rn(a, scale);
        const int64_t t0 = tok_offsets[idx], t1 = tok_offsets[idx + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) token_out[t] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// segment boundaries: src/prime_rl/trainer/utils.py:36-69.  Boundary at i iff pos[i]==0 && pos[i-1]!=0
// && i+1<T && pos[i+1]==1 (:54-60).  One CTA walks the row in 1024-token strips and appends the
// boundaries in order (ballot + warp-count prefix), so the reference's O(T) host/device syncs become
// one launch.  T <= a few 10^4: latency-bound by design, one CTA is the right size.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) segments_kernel(const int64_t* __restrict__ pos, int64_t T, int32_t* __restrict__ cu,
                                                        int32_t* __restrict__ nseg, int32_t max_seg) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    base = 1;
    cu[0] = 0;
  }
  __syncthreads();
  for (int64_t s0 = 0; s0 < T; s0 += 1024) {
    const int64_t i = s0 + threadIdx.x;
    bool b = false;
    if (i >= 1 && i + 1 < T) b = (pos[i] == 0) && (pos[i - 1] != 0) && (pos[i + 1] == 1);
    const unsigned bal = __ballot_sync(FULL, b);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < wid; ++w) off += warp_cnt[w];
    if (b) {
      const int slot = off + __popc(bal & ((1u << lane) - 1));
      if (slot <= max_seg - 1) cu[slot] = (int32_t)i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n = base;  // number of segments
    if (n <= max_seg) {
      cu[n] = (int32_t)T;
      *nseg = n;
    } else {
      *nseg = -n;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// packed loss.  src/prime_rl/trainer/rl/loss.py:61-124.
// token mode (:98-104): flat grid-stride pass, 29 B/token, per-CTA partial sums reduced in a fixed order.
// sequence mode (:92-96,107-108): one CTA per segment, two passes over the segment (second hits L1/L2).
// dlogp is what autograd would produce (clamp(max=) passes gradient at equality).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) loss_token_kernel(const float* __restrict__ logp, const float* __restrict__ old,
                                                         const float* __restrict__ adv, const uint8_t* __restrict__ mask, int64_t T,
                                                         float clip, float inv_scale, float* __restrict__ ratio,
                                                         float* __restrict__ clipped, float* __restrict__ is_clipped,
                                                         float* __restrict__ dlogp, double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t T4 = T / 4;
  const bool vec_ok = ((((uintptr_t)logp | (uintptr_t)old | (uintptr_t)adv | (uintptr_t)ratio | (uintptr_t)clipped |
                         (uintptr_t)is_clipped | (uintptr_t)dlogp) &
                        15) == 0) &&
                      (((uintptr_t)mask & 3) == 0);
  int64_t done = 0;
  if (vec_ok) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < T4; q += stride) {
      const float4 lp = reinterpret_cast<const float4*>(logp)[q];
      const float4 ol = reinterpret_cast<const float4*>(old)[q];
      const float4 av = reinterpret_cast<const float4*>(adv)[q];
      const uchar4 mk = reinterpret_cast<const uchar4*>(mask)[q];
      const float l[4] = {lp.x, lp.y, lp.z, lp.w}, o[4] = {ol.x, ol.y, ol.z, ol.w}, a[4] = {av.x, av.y, av.z, av.w};
      const bool m[4] = {mk.x != 0, mk.y != 0, mk.z != 0, mk.w != 0};
      float r[4], c[4], f[4], g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[k] = expf(l[k] - o[k]);
        c[k] = clamp_max_nan(r[k], clip);
        f[k] = (r[k] > clip) ? 1.f : 0.f;
        g[k] = (m[k] && r[k] <= clip) ? (-a[k] * r[k]) * inv_scale : 0.f;
        if (r[k] != r[k]) g[k] = r[k];  // autograd: 0 * exp'(nan) = nan, masked or not
        if (m[k]) acc += (double)(-c[k] * a[k]);
      }
      reinterpret_cast<float4*>(ratio)[q] = make_float4(r[0], r[1], r[2], r[3]);
      reinterpret_cast<float4*>(clipped)[q] = make_float4(c[0], c[1], c[2], c[3]);
      reinterpret_cast<float4*>(is_clipped)[q] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(dlogp)[q] = make_float4(g[0], g[1], g[2], g[3]);
    }
    done = T4 * 4;
  }
  for (int64_t t = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += stride) {
    const float r = expf(logp[t] - old[t]);
    const float c = clamp_max_nan(r, clip);
    const bool m = mask[t] != 0;
    ratio[t] = r;
    clipped[t] = c;
    is_clipped[t] = (r > clip) ? 1.f : 0.f;
    dlogp[t] = (r != r) ? r : ((m && r <= clip) ? (-adv[t] * r) * inv_scale : 0.f);
    if (m) acc += (double)(-c * adv[t]);
  }
  const double tot = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(512) loss_seq_kernel(const float* __restrict__ logp, const float* __restrict__ old,
                                                       const float* __restrict__ adv, const uint8_t* __restrict__ mask,
                                                       const int32_t* __restrict__ cu, int length_norm, float clip, float inv_scale,
                                                       float* __restrict__ ratio, float* __restrict__ clipped,
                                                       float* __restrict__ is_clipped, float* __restrict__ dlogp,
                                                       double* __restrict__ partial) {
  __shared__ double sh[16];
  __shared__ float coef_sh;
  const int seg = blockIdx.x;
  const int t0 = cu[seg], t1 = cu[seg + 1];
  double sd = 0.0, sa = 0.0, sn = 0.0;
  auto take = [&](float l, float o, float a, bool m) {
    if (m) {
      sd += (double)(l - o);
      sa += (double)a;
      sn += 1.0;
    }
  };
  // 16-byte loads over the 4-aligned body of the segment (all four arrays share the index, so one alignment test on the
  // base pointers covers them); up to 3 scalar tokens at either end
  const bool vec_ok = ((((uintptr_t)logp | (uintptr_t)old | (uintptr_t)adv | (uintptr_t)dlogp) & 15) == 0) && (((uintptr_t)mask & 3) == 0);
  const int b0 = vec_ok ? min((t0 + 3) & ~3, t1) : t1;  // body = [b0, b1), both multiples of 4
  const int b1 = vec_ok ? max(b0, t1 & ~3) : t1;
  for (int t = t0 + threadIdx.x; t < b0; t += blockDim.x) take(logp[t], old[t], adv[t], mask[t] != 0);
  for (int q = b0 / 4 + threadIdx.x; q < b1 / 4; q += blockDim.x) {
    const float4 l = reinterpret_cast<const float4*>(logp)[q], o = reinterpret_cast<const float4*>(old)[q];
    const float4 a = reinterpret_cast<const float4*>(adv)[q];
    const uchar4 m = reinterpret_cast<const uchar4*>(mask)[q];
    take(l.x, o.x, a.x, m.x != 0);
    take(l.y, o.y, a.y, m.y != 0);
    take(l.z, o.z, a.z, m.z != 0);
    take(l.w, o.w, a.w, m.w != 0);
  }
  for (int t = b1 + threadIdx.x; t < t1; t += blockDim.x) take(logp[t], old[t], adv[t], mask[t] != 0);
  const double Sd = block_sum(sd, sh);
  const double Sa = block_sum(sa, sh);
  const double Sn = block_sum(sn, sh);
  if (threadIdx.x == 0) {
    const float n = fmaxf((float)Sn, 1.f);
    float L = (float)Sd;
    if (length_norm) L = L / n;
    const float pre = L;
    L = clamp_max_nan(L, 10.f);
    const float r = expf(L);
    const float c = clamp_max_nan(r, clip);
    ratio[seg] = r;
    clipped[seg] = c;
    is_clipped[seg] = (r > clip) ? 1.f : 0.f;
    const float mean_a = (float)Sa / n;
    partial[seg] = (double)(-c * mean_a);
    float coef = -mean_a * r;
    if (!(r <= clip) || !(pre <= 10.f)) coef = 0.f;  // also for a NaN segment: clamp's backward is where(x <= max, g, 0), not g * mask
    if (length_norm) coef = coef / n;
    coef_sh = coef * inv_scale;
  }
  __syncthreads();
  const float coef = coef_sh;
  for (int t = t0 + threadIdx.x; t < b0; t += blockDim.x) dlogp[t] = mask[t] ? coef : 0.f;
  for (int q = b0 / 4 + threadIdx.x; q < b1 / 4; q += blockDim.x) {
    const uchar4 m = reinterpret_cast<const uchar4*>(mask)[q];
    reinterpret_cast<float4*>(dlogp)[q] = make_float4(m.x ? coef : 0.f, m.y ? coef : 0.f, m.z ? coef : 0.f, m.w ? coef : 0.f);
  }
  for (int t = b1 + threadIdx.x; t < t1; t += blockDim.x) dlogp[t] = mask[t] ? coef : 0.f;
}

// Sums partial[0..n) in fp64 in a FIXED order and scales: deterministic, and for n <= 32 (every trainer shape: one partial per
// segment or per CTA of a 16k-token row) exactly the left-to-right order in which the reference adds the segment losses
// (loss.py:110).  Longer lists (bench sizes: thousands of partials) are summed as 32 strided lane sums combined by a fixed
// shuffle tree -- one dependent chain of n global loads in one thread cost 50-75 us there.
__global__ void __launch_bounds__(32) loss_finalize_kernel(const double* __restrict__ partial, int n, double inv_scale, float* __restrict__ loss) {
  const int lane = threadIdx.x;
  double t = 0.0;
  if (n <= 32) {
    if (lane == 0)
      for (int i = 0; i < n; ++i) t += partial[i];
  } else {
    for (int i = lane; i < n; i += 32) t += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(FULL, t, o);
  }
  if (lane == 0) *loss = (float)(t * inv_scale);
}

__global__ void scale_kernel(const float* __restrict__ a, const float* __restrict__ scale, float* __restrict__ out, int64_t n) {
  const float s = *scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] * s;
}

// ------------------------------------------------------------------------------------------------
// API-compat statistics on materialised fp32 logits: one CTA per row, one pass (4 B/logit), online
// (max, sum-exp, sum exp*z).  rl/loss.py:26-30 (logp), :43-47 (entropy), :51-58 (shift).
// ------------------------------------------------------------------------------------------------
struct Stat {
  float m, s, u;
};
__device__ __forceinline__ Stat stat_merge(Stat a, Stat b) {
  const float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return a;
  const float ea = __expf(a.m - m), eb = __expf(b.m - m);
  Stat r;
  r.m = m;
  r.s = a.s * ea + b.s * eb;
  r.u = a.u * ea + b.u * eb;
  return r;
}
__device__ __forceinline__ void stat_push(Stat& st, float z) {
  if (z == -INFINITY) {
    // exp(-inf) = 0: nothing for the log-sum-exp (torch.logsumexp handles -inf logits), but the reference's entropy has the
    // term softmax * z = 0 * -inf = nan (rl/loss.py:45-46)
    st.u += __int_as_float(0x7fc00000);
    return;
  }
  if (z > st.m) {
    const float e = __expf(st.m - z);
    st.s = st.s * e + 1.f;
    st.u = st.u * e + z;
    st.m = z;
  } else {
    const float e = __expf(z - st.m);
    st.s += e;
    st.u += e * z;
  }
}

__global__ void __launch_bounds__(256) logits_stats_kernel(const float* __restrict__ logits, const int64_t* __restrict__ ids,
                                                           int64_t rows, int64_t seq, int64_t V, float inv_temp, int shift,
                                                           float* __restrict__ logp, float* __restrict__ entropy) {
  __shared__ Stat sh[8];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int64_t p = row % seq;
    if (shift && p == 0) {
      if (threadIdx.x == 0) {
        const float l = logf((float)V);
        if (logp) logp[row] = -l;
        if (entropy) entropy[row] = l;
      }
      continue;
    }
    const float* src = logits + (row - (shift ? 1 : 0)) * V;
    Stat st = {-INFINITY, 0.f, 0.f};
    if ((((uintptr_t)src) & 15) == 0) {
      const int64_t V4 = V / 4;
      for (int64_t q = threadIdx.x; q < V4; q += blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(src)[q];
        stat_push(st, x.x * inv_temp);
        stat_push(st, x.y * inv_temp);
        stat_push(st, x.z * inv_temp);
        stat_push(st, x.w * inv_temp);
      }
      for (int64_t v = V4 * 4 + threadIdx.x; v < V; v += blockDim.x) stat_push(st, src[v] * inv_temp);
    } else {
      for (int64_t v = threadIdx.x; v < V; v += blockDim.x) stat_push(st, src[v] * inv_temp);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Stat other;
      other.m = __shfl_down_sync(FULL, st.m, o);
      other.s = __shfl_down_sync(FULL, st.s, o);
      other.u = __shfl_down_sync(FULL, st.u, o);
      st = stat_merge(st, other);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
      Stat t = sh[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = stat_merge(t, sh[w]);
      const float lse = t.m + logf(t.s);
      if (logp && ids) logp[row] = src[ids[row]] * inv_temp - lse;
      if (entropy) entropy[row] = lse - t.u / t.s;
    }
  }
}

// backward of the same pass w.r.t. the materialised logits (autograd of rl/loss.py:26-30 as the reference trainer runs it):
//   dlogits[q, v] = g[r] / tau * (1[v == ids[r]] - exp(logits[q, v] / tau - lse[r])),  r = q (+1 if shift; none for the last position)
// lse is rebuilt from the forward's logp: lse[r] = logits[q, ids[r]] / tau - logp[r].  One read + one write of [rows, V] fp32.
__global__ void __launch_bounds__(256) logits_logsoftmax_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ ids,
                                                                    const float* __restrict__ logp, const float* __restrict__ g,
                                                                    int64_t rows, int64_t seq, int64_t V, float inv_temp, int shift,
                                                                    float* __restrict__ dlogits) {
  for (int64_t q = blockIdx.x; q < rows; q += gridDim.x) {
    const float* src = logits + q * V;
    float* dst = dlogits + q * V;
    const bool vec = ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0;
    const int64_t V4 = vec ? V / 4 : 0;
    const bool has = !shift || (q % seq) + 1 < seq;
    if (!has) {  // the position shift_logits drops (rl/loss.py:54): no gradient
      for (int64_t k = threadIdx.x; k < V4; k += blockDim.x) reinterpret_cast<float4*>(dst)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t v = V4 * 4 + threadIdx.x; v < V; v += blockDim.x) dst[v] = 0.f;
      continue;
    }
    const int64_t r = q + (shift ? 1 : 0);
    const int64_t id = ids[r];
    const float gr = g[r] * inv_temp;
    const float lse = src[id] * inv_temp - logp[r];
    for (int64_t k = threadIdx.x; k < V4; k += blockDim.x) {
      const float4 x = reinterpret_cast<const float4*>(src)[k];
      float4 d;
      d.x = -gr * __expf(x.x * inv_temp - lse);
      d.y = -gr * __expf(x.y * inv_temp - lse);
      d.z = -gr * __expf(x.z * inv_temp - lse);
      d.w = -gr * __expf(x.w * inv_temp - lse);
      const int64_t off = id - 4 * k;
      if (off == 0) d.x += gr;
      if (off == 1) d.y += gr;
      if (off == 2) d.z += gr;
      if (off == 3) d.w += gr;
      reinterpret_cast<float4*>(dst)[k] = d;
    }
    for (int64_t v = V4 * 4 + threadIdx.x; v < V; v += blockDim.x) dst[v] = -gr * __expf(src[v] * inv_temp - lse) + (v == id ? gr : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// masked running moments for logging (train.py:313-333 + trainer/utils.py:228-256): 5 B/element read, fp64 sums,
// fixed-order final combine (deterministic), accumulator updated in place.
// ------------------------------------------------------------------------------------------------
constexpr int MOM_BLOCKS = 148 * 4;

__global__ void __launch_bounds__(256) masked_moments_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n,
                                                             int transform, double* __restrict__ partial) {
  __shared__ double sh[5][8];
  double cnt = 0.0, sum = 0.0, sq = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask == nullptr || mask[i]) {
      float v = x[i];
      if (transform == 1) v = expf(v);
      const double d = (double)v;
      cnt += 1.0;
      sum += d;
      sq += d * d;
      mn = min_nan(mn, d);
      mx = max_nan(mx, d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_down_sync(FULL, cnt, o);
    sum += __shfl_down_sync(FULL, sum, o);
    sq += __shfl_down_sync(FULL, sq, o);
    mn = min_nan(mn, __shfl_down_sync(FULL, mn, o));
    mx = max_nan(mx, __shfl_down_sync(FULL, mx, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][wid] = cnt, sh[1][wid] = sum, sh[2][wid] = sq, sh[3][wid] = mn, sh[4][wid] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      cnt += sh[0][w], sum += sh[1][w], sq += sh[2][w];
      mn = min_nan(mn, sh[3][w]), mx = max_nan(mx, sh[4][w]);
    }
    double* o = partial + (int64_t)blockIdx.x * 5;
    o[0] = cnt, o[1] = sum, o[2] = sq, o[3] = mn, o[4] = mx;
  }
}

// ------------------------------------------------------------------------------------------------
// exact median of the logged values (N1): order-preserving u32 keys + 3-pass radix select (11 + 11 + 10 bits), so that ranks
// exchange three 2048-bin histograms per step instead of every masked token (trainer/utils.py:241-247)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// appends key(f(x[i])) for every unmasked i to keys[*counter ...] (warp-aggregated atomic; order is irrelevant for a selection)
__global__ void __launch_bounds__(256) masked_keys_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n,
                                                          int transform, uint32_t* __restrict__ keys, int64_t capacity,
                                                          unsigned long long* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) / 32 * 32;  // whole warps take part in every ballot
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool take = i < n && (mask == nullptr || mask[i]);
    const unsigned ballot = __ballot_sync(FULL, take);
    if (ballot == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
    base = __shfl_sync(FULL, base, 0);
    if (take) {
      float v = x[i];
      if (transform == 1) v = expf(v);
      const int64_t pos = (int64_t)base + __popc(ballot & ((1u << lane) - 1u));
      if (pos < capacity) keys[pos] = float_order_key(v);
    }
  }
}

// state = {prefix, prefix_mask, k, unused}: histogram of bits [shift, shift + bits) over the keys whose masked bits equal prefix
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t* __restrict__ keys, const unsigned long long* __restrict__ counter,
                                                         int64_t capacity, const int64_t* __restrict__ state, int shift, int bits,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[2048];
  for (int b = threadIdx.x; b < 2048; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  int64_t n = (int64_t)*counter;
  if (n > capacity) n = capacity;
  const uint32_t prefix = (uint32_t)state[0], pmask = (uint32_t)state[1], bmask = (1u << bits) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if ((k & pmask) == prefix) atomicAdd(&sh[(k >> shift) & bmask], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 2048; b += blockDim.x)
    if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
}

// one warp: find the bin holding the k-th (0-based) remaining element, extend the prefix, clear the histogram for the next pass
__global__ void radix_pick_kernel(unsigned long long* __restrict__ hist, int64_t* __restrict__ state, int shift, int bits) {
  const int lane = threadIdx.x;
  const int nb = 1 << bits, per = nb / 32;
  int64_t k = state[2];
  unsigned long long mine = 0;
  for (int b = lane * per; b < (lane + 1) * per; ++b) mine += hist[b];
  unsigned long long incl = mine;  // inclusive scan over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned long long excl = incl - mine;
  const bool here = k >= 0 && (unsigned long long)k >= excl && (unsigned long long)k < incl;
  const unsigned ballot = __ballot_sync(FULL, here);
  if (here) {
    unsigned long long cum = excl;
    for (int b = lane * per; b < (lane + 1) * per; ++b) {
      const unsigned long long c = hist[b];
      if ((unsigned long long)k < cum + c) {
        state[0] = (int64_t)((uint32_t)state[0] | ((uint32_t)b << shift));
        state[1] = (int64_t)((uint32_t)state[1] | (((1u << bits) - 1u) << shift));
        state[2] = k - (int64_t)cum;
        break;
      }
      cum += c;
    }
  }
  if (ballot == 0 && lane == 0) state[2] = -1;  // k out of range (no elements): stays unresolved
  __syncwarp();
  for (int b = lane; b < 2048; b += 32) hist[b] = 0;
}

__global__ void masked_moments_finalize_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ acc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double cnt = acc[0], sum = acc[1], sq = acc[2], mn = acc[3], mx = acc[4];
    for (int b = 0; b < nblocks; ++b) {
      const double* p = partial + (int64_t)b * 5;
      cnt += p[0], sum += p[1], sq += p[2];
      mn = min_nan(mn, p[3]), mx = max_nan(mx, p[4]);
    }
    acc[0] = cnt, acc[1] = sum, acc[2] = sq, acc[3] = mn, acc[4] = mx;
  }
}

// ------------------------------------------------------------------------------------------------
// rollouts -> packed rows (orchestrator/batch.py:21-62,105-120): one CTA per rollout, 13 B read + 25 B written per token
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_rollouts_kernel(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ mask,
                                                            const float* __restrict__ lp, const int64_t* __restrict__ src_tok_off,
                                                            const int64_t* __restrict__ src_lp_off, const int32_t* __restrict__ prompt_len,
                                                            const float* __restrict__ advantage, const int64_t* __restrict__ dst_off,
                                                            int64_t n, int64_t* __restrict__ input_ids, int64_t* __restrict__ position_ids,
                                                            float* __restrict__ advantages, float* __restrict__ logprobs,
                                                            uint8_t* __restrict__ loss_mask) {
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const int64_t s0 = src_tok_off[i], len = src_tok_off[i + 1] - s0, d0 = dst_off[i], l0 = src_lp_off[i];
    const int64_t pl = prompt_len[i];
    const float a = advantage[i];
    for (int64_t j = threadIdx.x; j < len; j += blockDim.x) {
      input_ids[d0 + j] = tokens[s0 + j];
      position_ids[d0 + j] = j;
      advantages[d0 + j] = a;
      logprobs[d0 + j] = (j < pl) ? 0.f : lp[l0 + j - pl];
      loss_mask[d0 + j] = mask[s0 + j] != 0 ? 1 : 0;
    }
  }
}

}  // namespace prl

// ================================================================================================
// C ABI
// ================================================================================================
using namespace prl;

extern "C" {

int prl_version(void) { return 100; }

const char* prl_strerror(int code) {
  switch (code) {
    case PRL_OK: return "ok";
    case PRL_ERR_BAD_ARG: return "bad argument";
    case PRL_ERR_UNSUPPORTED: return "unsupported shape or configuration";
    case PRL_ERR_CUDA: return "CUDA error";
    case PRL_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

const char* prl_last_error_detail(void) { return err_state().detail; }

int64_t prl_launch_count(int32_t reset) {
  const int64_t n = err_state().launches;
  if (reset) err_state().launches = 0;
  return n;
}

int prl_advantages_f32(const float* rewards, const int64_t* lengths, int64_t n, int32_t group, int32_t kind, float* out,
                       const int64_t* tok_offsets, float* token_out, void* stream) {
  PRL_REQUIRE(n >= 0 && group >= 1, PRL_ERR_BAD_ARG, "advantages: n >= 0 and group >= 1 required");
  PRL_REQUIRE(n % group == 0, PRL_ERR_BAD_ARG, "advantages: len(rewards) % samples_per_problem != 0");
  PRL_REQUIRE(kind >= 0 && kind <= 4, PRL_ERR_BAD_ARG, "advantages: unknown advantage kind");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(rewards && out, PRL_ERR_BAD_ARG, "advantages: null pointer");
  PRL_REQUIRE(kind != PRL_ADV_OPO || lengths, PRL_ERR_BAD_ARG, "advantages: opo needs completion lengths");
  PRL_REQUIRE((token_out == nullptr) || tok_offsets, PRL_ERR_BAD_ARG, "advantages: token_out needs tok_offsets");
  const int64_t ngroups = n / group;
  const int64_t blocks_needed = (ngroups + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(rewards) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(lengths)) & 15) == 0;
  if (token_out == nullptr && aligned16 && group % 8 == 0 && group <= 32) {
    const int m = group / 8, gpw = 32 / m;
    const int64_t need = ((ngroups + gpw - 1) / gpw + 7) / 8;  // one group-set per warp iteration, 8 warps per CTA
    const int g3 = (int)(need < cap * 2 ? need : cap * 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 1) advantages_vec8_kernel<1><<<g3, 256, 0, st>>>(rewards, lengths, ngroups, kind, out);
    if (m == 2) advantages_vec8_kernel<2><<<g3, 256, 0, st>>>(rewards, lengths, ngroups, kind, out);
    if (m == 3) advantages_vec8_kernel<3><<<g3, 256, 0, st>>>(rewards, lengths, ngroups, kind, out);
    if (m == 4) advantages_vec8_kernel<4><<<g3, 256, 0, st>>>(rewards, lengths, ngroups, kind, out);
    return check_launch("advantages_vec8_kernel");
  }
  if (group <= 32 && token_out == nullptr) {
    const int gpw = 32 / group;
    const int64_t need = (((ngroups + gpw - 1) / gpw + 3) / 4 + 7) / 8;  // 4 group-sets per warp iteration, 8 warps per CTA
    const int g2 = (int)(need < cap * 4 ? need : cap * 4);
    advantages_small_group_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(rewards, lengths, ngroups, group, kind, out);
    return check_launch("advantages_small_group_kernel");
  }
  advantages_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rewards, lengths, ngroups, group, kind, out, tok_offsets, token_out);
  return check_launch("advantages_kernel");
}

int prl_segments_from_position_ids(const int64_t* position_ids, int64_t T, int32_t* cu_seqlens, int32_t* nseg, int32_t max_seg,
                                   void* stream) {
  PRL_REQUIRE(position_ids && cu_seqlens && nseg, PRL_ERR_BAD_ARG, "segments: null pointer");
  PRL_REQUIRE(T >= 1 && T < (int64_t)1 << 31 && max_seg >= 1, PRL_ERR_BAD_ARG, "segments: 1 <= T < 2^31 and max_seg >= 1 required");
  segments_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(position_ids, T, cu_seqlens, nseg, max_seg);
  return check_launch("segments_kernel");
}

static inline int loss_token_grid(int64_t T) {
  const int64_t need = (T / 4 + 255) / 256 + 1;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(need < cap ? need : cap);
}

size_t prl_loss_workspace_bytes(int64_t T, int32_t nseg) {
  const int64_t a = (int64_t)148 * 8 * 2, b = nseg > 0 ? nseg : 1;
  (void)T;
  return sizeof(double) * (size_t)(a > b ? a : b);
}

int prl_loss_packed_fwd(const float* logp, const float* old_logp, const float* adv, const uint8_t* mask, const int32_t* cu_seqlens,
                        int32_t nseg, int64_t T, int32_t mode, int32_t length_norm, float clip_ratio, double loss_scale, float* loss,
                        float* ratio, float* clipped, float* is_clipped, float* dlogp, void* workspace, size_t workspace_bytes,
                        void* stream) {
  PRL_REQUIRE(T >= 0 && nseg >= 0, PRL_ERR_BAD_ARG, "loss: negative size");
  PRL_REQUIRE(loss && workspace, PRL_ERR_BAD_ARG, "loss: null loss/workspace pointer");
  PRL_REQUIRE(mode == PRL_RATIO_TOKEN || mode == PRL_RATIO_SEQUENCE, PRL_ERR_BAD_ARG, "loss: unknown ratio_type");
  PRL_REQUIRE(workspace_bytes >= prl_loss_workspace_bytes(T, nseg), PRL_ERR_WORKSPACE, "loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const double S = loss_scale > 1.0 ? loss_scale : 1.0;  // max(loss_scale, 1), loss.py:118
  double* partial = reinterpret_cast<double*>(workspace);
  if (T == 0) {
    loss_finalize_kernel<<<1, 32, 0, st>>>(partial, 0, 1.0 / S, loss);
    return check_launch("loss_finalize_kernel");
  }
  PRL_REQUIRE(logp && old_logp && adv && mask && ratio && clipped && is_clipped && dlogp, PRL_ERR_BAD_ARG, "loss: null pointer");
  int nparts;
  if (mode == PRL_RATIO_TOKEN) {
    nparts = loss_token_grid(T);
    loss_token_kernel<<<nparts, 256, 0, st>>>(logp, old_logp, adv, mask, T, clip_ratio, (float)(1.0 / S), ratio, clipped,
                                              is_clipped, dlogp, partial);
    int rc = check_launch("loss_token_kernel");
    if (rc) return rc;
  } else {
    PRL_REQUIRE(cu_seqlens && nseg >= 1, PRL_ERR_BAD_ARG, "loss: sequence mode needs cu_seqlens and nseg >= 1");
    nparts = nseg;
    loss_seq_kernel<<<nseg, 512, 0, st>>>(logp, old_logp, adv, mask, cu_seqlens, length_norm, clip_ratio, (float)(1.0 / S), ratio,
                                          clipped, is_clipped, dlogp, partial);
    int rc = check_launch("loss_seq_kernel");
    if (rc) return rc;
  }
  loss_finalize_kernel<<<1, 32, 0, st>>>(partial, nparts, 1.0 / S, loss);
  return check_launch("loss_finalize_kernel");
}

int prl_scale_f32(const float* a, const float* scale, float* out, int64_t n, void* stream) {
  PRL_REQUIRE(n >= 0, PRL_ERR_BAD_ARG, "scale: negative size");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(a && scale && out, PRL_ERR_BAD_ARG, "scale: null pointer");
  const int64_t need = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  scale_kernel<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(a, scale, out, n);
  return check_launch("scale_kernel");
}

int prl_pack_rollouts(const int64_t* tokens, const uint8_t* mask, const float* completion_logprobs, const int64_t* src_tok_off,
                      const int64_t* src_lp_off, const int32_t* prompt_len, const float* advantage, const int64_t* dst_off, int64_t n,
                      int64_t* input_ids, int64_t* position_ids, float* advantages, float* logprobs, uint8_t* loss_mask, void* stream) {
  PRL_REQUIRE(n >= 0, PRL_ERR_BAD_ARG, "pack_rollouts: negative size");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(tokens && mask && completion_logprobs && src_tok_off && src_lp_off && prompt_len && advantage && dst_off && input_ids &&
                  position_ids && advantages && logprobs && loss_mask,
              PRL_ERR_BAD_ARG, "pack_rollouts: null pointer");
  const int64_t cap = (int64_t)sm_count() * 8;
  pack_rollouts_kernel<<<(int)(n < cap ? n : cap), 256, 0, (cudaStream_t)stream>>>(tokens, mask, completion_logprobs, src_tok_off, src_lp_off,
                                                                                    prompt_len, advantage, dst_off, n, input_ids, position_ids,
                                                                                    advantages, logprobs, loss_mask);
  return check_launch("pack_rollouts_kernel");
}

size_t prl_masked_moments_workspace_bytes(void) { return sizeof(double) * 5 * MOM_BLOCKS; }

int prl_masked_moments_f32(const float* x, const uint8_t* mask, int64_t n, int32_t transform, double* acc, void* workspace,
                           size_t workspace_bytes, void* stream) {
  PRL_REQUIRE(n >= 0, PRL_ERR_BAD_ARG, "masked_moments: negative size");
  PRL_REQUIRE(transform == 0 || transform == 1, PRL_ERR_BAD_ARG, "masked_moments: unknown transform");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(x && acc && workspace, PRL_ERR_BAD_ARG, "masked_moments: null pointer");
  PRL_REQUIRE(workspace_bytes >= prl_masked_moments_workspace_bytes(), PRL_ERR_WORKSPACE, "masked_moments: workspace too small");
  const int64_t need = (n + 255) / 256;
  const int grid = (int)(need < MOM_BLOCKS ? need : MOM_BLOCKS);
  cudaStream_t st = (cudaStream_t)stream;
  masked_moments_kernel<<<grid, 256, 0, st>>>(x, mask, n, transform, reinterpret_cast<double*>(workspace));
  int rc = check_launch("masked_moments_kernel");
  if (rc) return rc;
  masked_moments_finalize_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, acc);
  return check_launch("masked_moments_finalize_kernel");
}

int prl_masked_keys_u32(const float* x, const uint8_t* mask, int64_t n, int32_t transform, uint32_t* keys, int64_t capacity,
                        unsigned long long* counter, void* stream) {
  PRL_REQUIRE(n >= 0 && capacity >= 0, PRL_ERR_BAD_ARG, "masked_keys: negative size");
  PRL_REQUIRE(transform == 0 || transform == 1, PRL_ERR_BAD_ARG, "masked_keys: unknown transform");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(x && keys && counter, PRL_ERR_BAD_ARG, "masked_keys: null pointer");
  const int64_t need = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  masked_keys_kernel<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(x, mask, n, transform, keys, capacity, counter);
  return check_launch("masked_keys_kernel");
}

int prl_radix_select_hist(const uint32_t* keys, const unsigned long long* counter, int64_t capacity, const int64_t* state, int32_t shift,
                          int32_t bits, unsigned long long* hist, void* stream) {
  PRL_REQUIRE(capacity >= 0 && shift >= 0 && bits >= 1 && bits <= 11 && shift + bits <= 32, PRL_ERR_BAD_ARG, "radix_select_hist: bad sizes");
  if (capacity == 0) return PRL_OK;
  PRL_REQUIRE(keys && counter && state && hist, PRL_ERR_BAD_ARG, "radix_select_hist: null pointer");
  const int64_t need = (capacity + 255) / 256, cap = (int64_t)sm_count() * 4;
  radix_hist_kernel<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(keys, counter, capacity, state, shift, bits, hist);
  return check_launch("radix_hist_kernel");
}

int prl_radix_select_pick(unsigned long long* hist, int64_t* state, int32_t shift, int32_t bits, void* stream) {
  PRL_REQUIRE(shift >= 0 && bits >= 5 && bits <= 11 && shift + bits <= 32, PRL_ERR_BAD_ARG, "radix_select_pick: bad sizes");
  PRL_REQUIRE(hist && state, PRL_ERR_BAD_ARG, "radix_select_pick: null pointer");
  radix_pick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hist, state, shift, bits);
  return check_launch("radix_pick_kernel");
}

int prl_logits_stats_f32(const float* logits, const int64_t* ids, int64_t rows, int64_t seq, int64_t V, float inv_temp,
                         int32_t shift, float* logp, float* entropy, void* stream) {
  PRL_REQUIRE(rows >= 0 && seq >= 1 && V >= 1, PRL_ERR_BAD_ARG, "logits_stats: bad sizes");
  if (rows == 0) return PRL_OK;
  PRL_REQUIRE(logits, PRL_ERR_BAD_ARG, "logits_stats: null logits");
  PRL_REQUIRE(logp == nullptr || ids != nullptr, PRL_ERR_BAD_ARG, "logits_stats: logp needs ids");
  const int64_t cap = (int64_t)sm_count() * 16;
  logits_stats_kernel<<<(int)(rows < cap ? rows : cap), 256, 0, (cudaStream_t)stream>>>(logits, ids, rows, seq, V, inv_temp, shift,
                                                                                        logp, entropy);
  return check_launch("logits_stats_kernel");
}

int prl_logits_logsoftmax_bwd_f32(const float* logits, const int64_t* ids, const float* logp, const float* dlogp, int64_t rows,
                                  int64_t seq, int64_t V, float inv_temp, int32_t shift, float* dlogits, void* stream) {
  PRL_REQUIRE(rows >= 0 && seq >= 1 && V >= 1, PRL_ERR_BAD_ARG, "logits_logsoftmax_bwd: bad sizes");
  if (rows == 0) return PRL_OK;
  PRL_REQUIRE(logits && ids && logp && dlogp && dlogits, PRL_ERR_BAD_ARG, "logits_logsoftmax_bwd: null pointer");
  PRL_REQUIRE(logits != dlogits, PRL_ERR_BAD_ARG, "logits_logsoftmax_bwd: in-place operation is not supported");
  const int64_t cap = (int64_t)sm_count() * 16;
  logits_logsoftmax_bwd_kernel<<<(int)(rows < cap ? rows : cap), 256, 0, (cudaStream_t)stream>>>(logits, ids, logp, dlogp, rows, seq, V,
                                                                                                 inv_temp, shift, dlogits);
  return check_launch("logits_logsoftmax_bwd_kernel");
}

}  // extern "C"
