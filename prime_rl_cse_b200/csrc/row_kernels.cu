// Masked positions of a packed row (prompt / environment tokens): the reference multiplies their loss by 0
// (src/prime_rl/trainer/rl/loss.py:104, mask built in src/prime_rl/orchestrator/batch.py:41-44), so their dlogits rows are
// exactly zero and contribute nothing to d_hidden / d_W.  These kernels put the ACTIVE positions first so that the GEMM
// engine only visits them: a stable partition of the positions (one launch), a row gather of `hidden` into that order, the
// inverse scatter of the per-token outputs and of d_hidden.  All HBM-bound copies (2 x 2 x T x H bytes per micro-batch, ~0.1 %
// of a step); no host-side index tensors.
#include "common.cuh"

namespace prl {

constexpr unsigned FULLMASK = 0xffffffffu;

// Stable partition of positions [shift, T): positions with mask != 0 first (in order), then the others (in order).
//   order[j]  = position held by slot j of the partitioned row, j in [0, T - shift)
//   slot[t]   = slot of position t (t >= shift); slot[t] = -1 for t < shift (no source row: shift_logits' zero row)
//   *n_active = number of active positions
// One CTA walks the row in 1024-token strips twice (count, then place): T <= a few 10^4 => latency-bound, one launch.
__global__ void __launch_bounds__(1024) partition_positions_kernel(const uint8_t* __restrict__ mask, int64_t T, int32_t shift,
                                                                   int32_t* __restrict__ order, int32_t* __restrict__ slot,
                                                                   int32_t* __restrict__ n_active) {
  __shared__ int warp_cnt[32];
  __shared__ int base_act, base_in;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // pass 1: count
  int cnt = 0;
  for (int64_t i = shift + threadIdx.x; i < T; i += 1024) cnt += mask[i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(FULLMASK, cnt, o);
  if (lane == 0) warp_cnt[wid] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
    base_act = 0;
    base_in = tot;
    *n_active = tot;
  }
  for (int64_t t = threadIdx.x; t < shift && t < T; t += 1024) slot[t] = -1;
  __syncthreads();
  // pass 2: place (ordered ballot compaction, strip by strip)
  for (int64_t s0 = shift; s0 < T; s0 += 1024) {
    const int64_t i = s0 + threadIdx.x;
    const bool in = i < T;
    const bool act = in && mask[i] != 0;
    const unsigned bal = __ballot_sync(FULLMASK, act);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int act_before = 0;  // active tokens of this strip in earlier warps
    for (int w = 0; w < wid; ++w) act_before += warp_cnt[w];
    if (in) {
      const unsigned lower = (1u << lane) - 1;
      const int a = act_before + __popc(bal & lower);                              // active before me in this strip
      const int before = wid * 32 + lane;                                          // tokens before me in this strip
      const int s = act ? base_act + a : base_in + (before - a);
      order[s] = (int32_t)i;
      slot[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
      const int64_t n_strip = (T - s0) < 1024 ? (T - s0) : 1024;
      base_act += tot;
      base_in += (int)n_strip - tot;
    }
    __syncthreads();
  }
}

// dst[j, :] = src[order[j] - shift, :] (bf16 rows of H elements, 16-byte vectors) and ids_out[j] = ids[order[j]]
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ ids, const int32_t* __restrict__ order,
                                   int64_t n, int32_t shift, int32_t vec_per_row, uint4* __restrict__ dst,
                                   int64_t* __restrict__ ids_out) {
  const int64_t total = n * vec_per_row;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = q / vec_per_row;
    const int c = (int)(q - j * vec_per_row);
    const int64_t t = order[j];
    dst[q] = __ldg(src + (t - shift) * vec_per_row + c);
    if (c == 0 && ids_out) ids_out[j] = ids[t];
  }
}

// per-token fp32 vectors into slot order: out_k[j] = in_k[order[j]] for up to three vectors (lse, dlogp, ...)
__global__ void gather_tokens_kernel(const int32_t* __restrict__ order, int64_t n, const float* __restrict__ a, float* __restrict__ a_out,
                                     const float* __restrict__ b, float* __restrict__ b_out, const float* __restrict__ c,
                                     float* __restrict__ c_out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = order[j];
    if (a) a_out[j] = a[t];
    if (b) b_out[j] = b[t];
    if (c) c_out[j] = c[t];
  }
}

// back to position order: out_k[t] = (slot[t] in [0, n_valid)) ? in_k[slot[t]] : fill_k; positions t < shift (slot -1) get
// `first_k` (the closed form of shift_logits' all-zero row: logp = -ln V, entropy = lse = ln V)
__global__ void scatter_tokens_kernel(const int32_t* __restrict__ slot, int64_t T, int64_t n_valid, const float* __restrict__ a,
                                      float* __restrict__ a_out, float a_first, const float* __restrict__ b, float* __restrict__ b_out,
                                      float b_first, const float* __restrict__ c, float* __restrict__ c_out, float c_first) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    const int s = slot[t];
    const bool ok = s >= 0 && s < n_valid;
    if (a_out) a_out[t] = ok ? a[s] : (s < 0 ? a_first : 0.f);
    if (b_out) b_out[t] = ok ? b[s] : (s < 0 ? b_first : 0.f);
    if (c_out) c_out[t] = ok ? c[s] : (s < 0 ? c_first : 0.f);
  }
}

// d_hidden back to source-row order: dst[r, :] = slot[r + shift] in [0, n_valid) ? src[slot[r + shift], :] : 0, r in [0, T)
// (rows whose position is masked, and the last `shift` rows that feed no position, get an exact zero gradient)
__global__ void scatter_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ slot, int64_t T, int64_t n_valid,
                                    int32_t shift, int32_t vec_per_row, uint4* __restrict__ dst) {
  const int64_t total = T * vec_per_row;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / vec_per_row;
    const int c = (int)(q - r * vec_per_row);
    const int64_t t = r + shift;
    const int s = t < T ? slot[t] : -1;
    dst[q] = (s >= 0 && s < n_valid) ? __ldg(src + (int64_t)s * vec_per_row + c) : make_uint4(0u, 0u, 0u, 0u);
  }
}

static int grid_for(int64_t work_items, int threads) {
  const int64_t need = (work_items + threads - 1) / threads, cap = (int64_t)sm_count() * 16;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace prl

using namespace prl;

extern "C" {

int prl_partition_positions(const uint8_t* mask, int64_t T, int32_t shift, int32_t* order, int32_t* slot, int32_t* n_active,
                            void* stream) {
  PRL_REQUIRE(T >= 0 && T < (int64_t)1 << 31, PRL_ERR_BAD_ARG, "partition_positions: bad T");
  PRL_REQUIRE(shift == 0 || shift == 1, PRL_ERR_BAD_ARG, "partition_positions: shift must be 0 or 1");
  PRL_REQUIRE(n_active && (T == 0 || (mask && order && slot)), PRL_ERR_BAD_ARG, "partition_positions: null pointer");
  partition_positions_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mask, T, shift, order, slot, n_active);
  return check_launch("partition_positions_kernel");
}

int prl_gather_rows_bf16(const void* src, const int64_t* ids, const int32_t* order, int64_t n, int32_t shift, int32_t H, void* dst,
                         int64_t* ids_out, void* stream) {
  PRL_REQUIRE(n >= 0 && H >= 8 && H % 8 == 0, PRL_ERR_BAD_ARG, "gather_rows: bad sizes (H must be a multiple of 8)");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(src && order && dst && (!ids_out || ids), PRL_ERR_BAD_ARG, "gather_rows: null pointer");
  PRL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, PRL_ERR_BAD_ARG,
              "gather_rows: buffers must be 16-byte aligned");
  const int vpr = H / 8;
  gather_rows_kernel<<<grid_for(n * vpr, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), ids, order, n, shift, vpr,
                                                                               reinterpret_cast<uint4*>(dst), ids_out);
  return check_launch("gather_rows_kernel");
}

int prl_gather_tokens_f32(const int32_t* order, int64_t n, const float* a, float* a_out, const float* b, float* b_out, const float* c,
                          float* c_out, void* stream) {
  PRL_REQUIRE(n >= 0, PRL_ERR_BAD_ARG, "gather_tokens: bad n");
  if (n == 0) return PRL_OK;
  PRL_REQUIRE(order && (!a || a_out) && (!b || b_out) && (!c || c_out), PRL_ERR_BAD_ARG, "gather_tokens: null pointer");
  gather_tokens_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(order, n, a, a_out, b, b_out, c, c_out);
  return check_launch("gather_tokens_kernel");
}

int prl_scatter_tokens_f32(const int32_t* slot, int64_t T, int64_t n_valid, const float* a, float* a_out, float a_first, const float* b,
                           float* b_out, float b_first, const float* c, float* c_out, float c_first, void* stream) {
  PRL_REQUIRE(T >= 0 && n_valid >= 0, PRL_ERR_BAD_ARG, "scatter_tokens: bad sizes");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(slot && (!a_out || a || n_valid == 0) && (!b_out || b || n_valid == 0) && (!c_out || c || n_valid == 0), PRL_ERR_BAD_ARG,
              "scatter_tokens: null pointer");
  scatter_tokens_kernel<<<grid_for(T, 256), 256, 0, (cudaStream_t)stream>>>(slot, T, n_valid, a, a_out, a_first, b, b_out, b_first, c, c_out,
                                                                            c_first);
  return check_launch("scatter_tokens_kernel");
}

int prl_scatter_rows_bf16(const void* src, const int32_t* slot, int64_t T, int64_t n_valid, int32_t shift, int32_t H, void* dst,
                          void* stream) {
  PRL_REQUIRE(T >= 0 && n_valid >= 0 && H >= 8 && H % 8 == 0, PRL_ERR_BAD_ARG, "scatter_rows: bad sizes (H must be a multiple of 8)");
  if (T == 0) return PRL_OK;
  PRL_REQUIRE(slot && dst && (src || n_valid == 0), PRL_ERR_BAD_ARG, "scatter_rows: null pointer");
  PRL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, PRL_ERR_BAD_ARG,
              "scatter_rows: buffers must be 16-byte aligned");
  const int vpr = H / 8;
  scatter_rows_kernel<<<grid_for(T * vpr, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), slot, T, n_valid, shift, vpr,
                                                                                reinterpret_cast<uint4*>(dst));
  return check_launch("scatter_rows_kernel");
}

// ---- peer memory for the fused reduce-scatter (include/prl_b200.h): open a CUDA IPC handle exported by another rank of the
// same node under THIS rank's device, with peer access enabled lazily (no context on the exporting device is created here)
int prl_ipc_alloc(size_t bytes, void** dev_ptr_out, void* handle_out) {
  PRL_REQUIRE(bytes > 0 && dev_ptr_out && handle_out, PRL_ERR_BAD_ARG, "ipc_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail_cuda(e, "cudaMalloc (ipc_alloc)");
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    return fail_cuda(e, "cudaIpcGetMemHandle");
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr_out = p;
  return PRL_OK;
}

int prl_ipc_free(void* dev_ptr) {
  if (!dev_ptr) return PRL_OK;
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail_cuda(e, "cudaFree (ipc_free)");
  }
  return PRL_OK;
}

int prl_ipc_open(const void* handle, void** dev_ptr_out) {
  PRL_REQUIRE(handle && dev_ptr_out, PRL_ERR_BAD_ARG, "ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail_cuda(e, "cudaIpcOpenMemHandle");
  }
  *dev_ptr_out = p;
  return PRL_OK;
}

int prl_ipc_close(void* dev_ptr) {
  if (!dev_ptr) return PRL_OK;
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail_cuda(e, "cudaIpcCloseMemHandle");
  }
  return PRL_OK;
}

}  // extern "C"
