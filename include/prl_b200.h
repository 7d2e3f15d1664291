/*
 * prl_b200.h — C ABI of the B200-native PRIME-RL per-token RL loss hot path.
 *
 * The reference (pmahdavi/prime-rl-cse) is pure Python and has no FFI; every
 * entry point below replaces a Python-level interface of the reference, cited
 * as file:line relative to the reference root.  The Python host layer in
 * prime_rl_cse_b200/ binds these with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     stated otherwise; the caller owns all memory (inputs, outputs, workspace);
 *     the library never allocates, frees or retains device memory (one explicit
 *     exception: prl_ipc_alloc / prl_ipc_free, buffers shared between ranks).
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and
 *     the call returns without synchronising.
 *   - return value: PRL_OK (0) or a negative PRL_ERR_* code; nothing is thrown
 *     across the ABI.  prl_strerror() names the code, prl_last_error_detail()
 *     gives the thread-local detail string (e.g. the cudaGetErrorString text).
 *   - sm_100a only.  There is no CPU fallback: on a machine without a B200 every
 *     compute entry point returns PRL_ERR_CUDA.
 */
#ifndef PRL_B200_H_
#define PRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRL_OK 0
#define PRL_ERR_BAD_ARG (-1)     /* null pointer, negative size, n % group != 0 ...        */
#define PRL_ERR_UNSUPPORTED (-2) /* shape outside what the kernels handle (e.g. H % 8 != 0) */
#define PRL_ERR_CUDA (-3)        /* launch / driver error, see prl_last_error_detail()      */
#define PRL_ERR_WORKSPACE (-4)   /* workspace too small                                     */

/* advantage kinds: REGISTRY of src/prime_rl/orchestrator/advantage.py:63-68 */
#define PRL_ADV_DRGRPO 0
#define PRL_ADV_DRGRPO_NEGCLIPPED 1
#define PRL_ADV_RLOO 2
#define PRL_ADV_OPO 3
/* extension, NOT in the reference's registry (which is Dr.GRPO: mean only): GRPO's std-normalised advantage
 * (r - mean) / (std + PRL_ADV_GRPO_EPS), std unbiased (torch.std); BASELINE.json north_star names "per-group reward mean/std" */
#define PRL_ADV_GRPO 4
#define PRL_ADV_GRPO_EPS 1e-6f

/* ratio types: LossConfig.ratio_type, src/prime_rl/trainer/rl/config.py:22 */
#define PRL_RATIO_TOKEN 0
#define PRL_RATIO_SEQUENCE 1

/* dtype tags for gradient outputs */
#define PRL_DT_BF16 0
#define PRL_DT_F32 1

int prl_version(void);
const char* prl_strerror(int code);
const char* prl_last_error_detail(void);

/* ------------------------------------------------------------------------------------------
 * Group-relative advantages.
 * Replaces compute_advantages(), src/prime_rl/orchestrator/advantage.py:71-103 (per-group maths
 * :18, :32, :41-43, :55-57).  rewards[n] fp32, lengths[n] int64 (read only for PRL_ADV_OPO, may
 * be NULL otherwise), n % group == 0.  out[n] fp32.  One warp per group; the fp32 sum uses 8
 * strided lane accumulators combined left to right, then the tail (DESIGN.md "advantages").
 * If token_out != NULL also broadcasts out[i] over the tokens of rollout i:
 *   token_out[t] = out[i] for tok_offsets[i] <= t < tok_offsets[i+1]
 * (replaces prepare_sample's `.repeat(len)`, src/prime_rl/orchestrator/batch.py:44).
 * ---------------------------------------------------------------------------------------- */
int prl_advantages_f32(const float* rewards, const int64_t* lengths, int64_t n, int32_t group, int32_t kind,
                       float* out, const int64_t* tok_offsets, float* token_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Segment boundaries of a packed row.
 * Replaces get_response_lengths(), src/prime_rl/trainer/utils.py:36-69 (boundary rule :54-60).
 * position_ids[T] int64.  Writes cu_seqlens[0..nseg] (int32, cu_seqlens[0]=0, cu_seqlens[nseg]=T)
 * and *nseg (int32, device).  max_seg bounds the segment count that will be written; if more
 * boundaries exist *nseg is set to -(true count) and only max_seg are written.
 * ---------------------------------------------------------------------------------------- */
int prl_segments_from_position_ids(const int64_t* position_ids, int64_t T, int32_t* cu_seqlens, int32_t* nseg,
                                   int32_t max_seg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Packed clipped importance-ratio loss, forward (also emits d loss / d logp).
 * Replaces compute_loss(), src/prime_rl/trainer/rl/loss.py:61-124, on the packed layout of
 * MicroBatch (src/prime_rl/trainer/rl/data.py:13-23): logp/old/adv fp32 [T], mask uint8 [T]
 * (torch.bool storage), cu_seqlens int32 [nseg+1] (device).
 *   mode PRL_RATIO_TOKEN:    ratio/clipped/is_clipped are [T];
 *   mode PRL_RATIO_SEQUENCE: they are [nseg] (loss.py:92-96,107-108).
 * loss_scale is the caller's `loss_scale` (max(.,1) applied inside, loss.py:118).
 * loss: 1 float.  dlogp[T]: d loss / d logp (autograd equivalent).
 * workspace: prl_loss_workspace_bytes(T, nseg) bytes.
 * ---------------------------------------------------------------------------------------- */
size_t prl_loss_workspace_bytes(int64_t T, int32_t nseg);
int prl_loss_packed_fwd(const float* logp, const float* old_logp, const float* adv, const uint8_t* mask,
                        const int32_t* cu_seqlens, int32_t nseg, int64_t T, int32_t mode, int32_t length_norm,
                        float clip_ratio, double loss_scale, float* loss, float* ratio, float* clipped,
                        float* is_clipped, float* dlogp, void* workspace, size_t workspace_bytes, void* stream);

/* out[i] = a[i] * (*scale)   (backward of the scalar loss: dlogp * grad_output; scale is a device scalar) */
int prl_scale_f32(const float* a, const float* scale, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Unfused API-compat kernel on MATERIALISED logits: one pass over logits[R,V] fp32 giving
 *   logp[r] = z[r, ids[r]] - logsumexp(z[r]),  entropy[r] = logsumexp(z[r]) - sum softmax(z[r]) z[r]
 * with z = logits * inv_temp.  Replaces selective_log_softmax() (rl/loss.py:13-39, fp32 branch
 * :26-30) and compute_entropy() (rl/loss.py:43-47).  ids may be NULL (entropy only); logp / entropy
 * may be NULL.  If shift != 0 row r reads logits row r-1 and row 0 is all zeros (shift_logits,
 * rl/loss.py:51-58, applied per batch row of `seq` positions).
 * ---------------------------------------------------------------------------------------- */
int prl_logits_stats_f32(const float* logits, const int64_t* ids, int64_t rows, int64_t seq, int64_t V,
                         float inv_temp, int32_t shift, float* logp, float* entropy, void* stream);

/* Backward of the pass above w.r.t. the materialised logits (what autograd does behind rl/loss.py:26-30 when the
 * reference trainer is kept unchanged): dlogits[q, v] = dlogp[r] * inv_temp * (1[v == ids[r]] - softmax(z[r])_v), with
 * r = q for shift == 0, r = q + 1 (and zero for the last position of each `seq`-long row) for shift != 0.
 * logp[rows] is the forward's output (the row's logsumexp is rebuilt from it).  One read and one write of [rows, V]. */
int prl_logits_logsoftmax_bwd_f32(const float* logits, const int64_t* ids, const float* logp, const float* dlogp,
                                  int64_t rows, int64_t seq, int64_t V, float inv_temp, int32_t shift, float* dlogits,
                                  void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused LM head: hidden[T,H] x W[V,H]^T with an online (max, sum-exp, sum exp*z, target) epilogue;
 * the [T,V] logits never reach HBM.  Replaces the chain
 *   forward(model).logits  (src/prime_rl/trainer/model.py:250-254, HF lm_head nn.Linear)
 *   .float().contiguous()  (src/prime_rl/trainer/rl/train.py:285)
 *   shift_logits           (rl/loss.py:51-58)         [shift != 0]
 *   / temperature          (rl/train.py:287)          [inv_temp = 1/temperature]
 *   selective_log_softmax  (rl/loss.py:13-39)         -> logp[T]
 *   compute_entropy        (rl/loss.py:43-47)         -> entropy[T]
 * hidden, W: bf16 row-major (H contiguous), H % 8 == 0.  ids[T] int64 are GLOBAL vocab ids;
 * W holds vocab rows [v_offset, v_offset + V) of a vocabulary of V_total rows (V_total == V and
 * v_offset == 0 when not vocab-sharded).  round_bf16 != 0 rounds each accumulator to bf16 before the
 * softmax statistics (the reference's bf16 logits).
 * Outputs (each may be NULL): logp[T], entropy[T], lse[T] fp32 — final values, only meaningful when
 * the shard is the whole vocabulary; partials[T,4] fp32 = (m, s, u, z_target) of THIS shard, to be
 * exchanged across vocab shards and merged with prl_lmhead_merge_partials.
 * ---------------------------------------------------------------------------------------- */
size_t prl_lmhead_fwd_workspace_bytes(int64_t T, int32_t H, int64_t V);
int prl_lmhead_fwd(const void* hidden, const void* weight, const int64_t* ids, int64_t T, int32_t H, int64_t V,
                   int64_t v_offset, int64_t V_total, float inv_temp, int32_t shift, int32_t round_bf16,
                   float* logp, float* entropy, float* lse, float* partials, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Merge nparts per-token partials (partials[nparts, T, 4]) into logp / entropy / lse (vocab-sharded
 * LM head: the only data exchanged between vocab shards, BASELINE.json config 4). */
int prl_lmhead_merge_partials(const float* partials, int32_t nparts, int64_t T, float* logp, float* entropy,
                              float* lse, void* stream);

/* ------------------------------------------------------------------------------------------
 * Recompute-based backward of the fused LM head (autograd through rl/train.py:285-288, executed by
 * loss.backward() at rl/train.py:310).  Given lse[T] (from the forward) and dlogp[T] = d loss / d logp:
 *   dz[t,v] = dlogp[t] * (1[v == ids[t]] - exp(z[t,v] - lse[t])),  dlogits[t-shift] = dz[t] * inv_temp
 * (rounded to bf16 like the reference's bf16 logits gradient), then
 *   d_hidden[T,H] (bf16)            = dlogits x W
 *   d_weight[V,H] (bf16 or fp32)  (+)= dlogits^T x hidden      (accumulate_dw != 0 adds to d_weight)
 * Logits are recomputed tile by tile with the forward's GEMM; a bounded [T, vchunk] bf16 dlogits
 * block lives in `workspace` (never the whole [T,V]); vchunk == 0 picks the default.
 * d_hidden may be NULL (W-only) and d_weight may be NULL (hidden-only).  For a vocab shard d_hidden
 * is the shard's partial sum (reduce across shards by the caller); set d_hidden_f32 != 0 to get it as fp32.
 * phase_mask selects the sub-steps to run per vocab chunk (profiling aid): bit 0 = dlogits recompute,
 * bit 1 = d_hidden GEMM, bit 2 = d_weight GEMM; 0 means all three.
 * ---------------------------------------------------------------------------------------- */
size_t prl_lmhead_bwd_workspace_bytes(int64_t T, int32_t H, int64_t V, int64_t vchunk);
int prl_lmhead_bwd(const void* hidden, const void* weight, const int64_t* ids, const float* lse,
                   const float* dlogp, int64_t T, int32_t H, int64_t V, int64_t v_offset, float inv_temp,
                   int32_t shift, int32_t round_bf16, void* d_hidden, int32_t d_hidden_f32, void* d_weight,
                   int32_t d_weight_dtype, int32_t accumulate_dw, int64_t vchunk, int32_t phase_mask,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Vocab-sharded LM head (BASELINE.json config 4), backward with the reduce-scatter of d_hidden FUSED into the GEMM epilogue.
 * hidden / ids / lse / dlogp hold T = n_peers * rows_per_peer tokens: the packed rows of all peers back to back (token-owner
 * major), `weight` is this rank's vocab shard.  Instead of a local d_hidden, every output tile of the d_hidden GEMM is
 * reduce-added (fp32, cp.reduce.async.bulk.tensor performed at the owner's L2, over NVLink for remote owners) into
 * d_hidden_peers[owner][row - owner * rows_per_peer, :H], owner = row / rows_per_peer.  d_hidden_peers is a HOST array of
 * n_peers device pointers to fp32 [rows_per_peer, H] accumulators that are mapped into this process (CUDA IPC / symmetric
 * memory; entry self_index is this rank's own buffer; rows_per_peer must be a multiple of 32, so that a 32-row epilogue slab never
 * straddles two owners: PRL_ERR_UNSUPPORTED otherwise).  The caller zeroes its accumulator and synchronises the peers before
 * the call, and again before reading the result.  Replaces the NCCL reduce-scatter that follows the
 * GEMM in the unfused path; everything else (dlogits recompute, d_weight shard) is as in prl_lmhead_bwd. */
int prl_lmhead_bwd_scatter(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp,
                           int64_t T, int32_t H, int64_t V, int64_t v_offset, float inv_temp, int32_t shift, int32_t round_bf16,
                           void* const* d_hidden_peers, int32_t n_peers, int64_t rows_per_peer, int32_t self_index,
                           void* d_weight, int32_t d_weight_dtype, int32_t accumulate_dw, int64_t vchunk, int32_t phase_mask,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Token-sharded ranks (= FSDP ranks), gradient accumulation of d_W_lm straight INTO the reduce-scattered layout: instead of a
 * local fp32 [V, H] accumulator that FSDP2 / NCCL reduce-scatters at the end of the step (src/prime_rl/trainer/model.py:100-135,
 * reduce_dtype fp32), every tile of every micro-batch's d_weight GEMM is reduce-added (fp32, TMA, performed at the owner's L2,
 * over NVLink for remote owners) into the shard of the rank that owns those vocab rows: d_weight_shards[j] is rank j's fp32
 * [min(rows_per_peer, V - j * rows_per_peer), H] shard (rows_per_peer = ceil(V / n_peers): torch.chunk / FSDP2 dim-0 layout),
 * mapped into this process (prl_ipc_*); a 32-row slab that straddles two owners is sent to both with the row range clipped by
 * their tensor maps.  Ranks start at different vocab chunks so that they add into different owners at any time.  No step-end
 * collective is left (a barrier before the shards are read), and no rank holds a full [V, H] gradient.  The order in which the
 * ranks' addends reach a shard is not fixed: results are reproducible to fp32 rounding, not bit for bit.  d_hidden as in
 * prl_lmhead_bwd (local).  The caller zeroes its shard and synchronises the ranks before the first scattering micro-batch.
 * local_accumulator (fp32 [V, H], may be NULL): added tile by tile to what is sent -- so a trainer can accumulate the first
 * n - 1 micro-batches of a step locally (prl_lmhead_bwd, accumulate_dw) and let the LAST micro-batch's d_weight GEMM send
 * (local total + its own tiles): the NVLink traffic of one reduce-scatter per step, hidden behind that micro-batch's GEMMs. */
int prl_lmhead_bwd_dw_scatter(const void* hidden, const void* weight, const int64_t* ids, const float* lse, const float* dlogp,
                              int64_t T, int32_t H, int64_t V, float inv_temp, int32_t shift, int32_t round_bf16, void* d_hidden,
                              int32_t d_hidden_f32, void* const* d_weight_shards, int32_t n_peers, int64_t rows_per_peer,
                              int32_t self_index, const float* local_accumulator, int64_t vchunk, int32_t phase_mask,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Peer memory for prl_lmhead_bwd_scatter (the ONLY entry points that own device memory: an accumulator that other processes
 * map has to be a whole allocation of its own).  prl_ipc_alloc: cudaMalloc `bytes` on the current device and export it
 * (handle_out: 64 bytes, a cudaIpcMemHandle_t to hand to the other ranks of the node); prl_ipc_free releases it.
 * prl_ipc_open: map a buffer another rank exported into this process under the CURRENT device, enabling peer access lazily
 * (*dev_ptr_out = the exporter's allocation); prl_ipc_close unmaps it.  HOST pointers in, device pointers out. */
int prl_ipc_alloc(size_t bytes, void** dev_ptr_out, void* handle_out);
int prl_ipc_free(void* dev_ptr);
int prl_ipc_open(const void* handle, void** dev_ptr_out);
int prl_ipc_close(void* dev_ptr);

/* ------------------------------------------------------------------------------------------
 * Saved-logits variant of the same forward / backward pair: trades T x saved_ld x 2 bytes of HBM for the
 * backward's logits recompute (one of the four [T,H]x[H,V]-sized GEMMs of a micro-step).  Still replaces
 * rl/train.py:285-288,302-303,310; the reference itself keeps the fp32 [T,V] logits (and their gradient)
 * alive between forward and backward, this keeps ONE bf16 copy and overwrites it in place.
 *   prl_lmhead_saved_logits_ld(V): row pitch (elements) the library recommends for the saved block.
 *   prl_lmhead_fwd_save: prl_lmhead_fwd with round_bf16 = 1 that also writes the bf16 logits (before the
 *     1/temperature scaling) of source row r to saved_logits[r * saved_ld + v], r < T - shift, v < V.
 *   prl_lmhead_bwd_saved: converts saved_logits in place to dlogits (bit-identical to what the recompute
 *     backward produces; the block is CONSUMED), then runs the d_hidden / d_weight GEMMs over the whole
 *     vocabulary in one launch each (no vocab chunks).
 *     phase_mask: bit 0 = logits -> dlogits pass, bit 1 = d_hidden GEMM, bit 2 = d_weight GEMM; 0 = all.
 *     workspace (optional, prl_lmhead_bwd_saved_workspace_bytes; 0 bytes at the large shapes): an fp32 [T, H]
 *     accumulator that lets the d_hidden GEMM split K in two when T x H gives too few output tiles to fill
 *     the 148 SMs; without it the GEMM runs unsplit (same result up to fp32 summation order).
 * saved_logits: 16-byte aligned, saved_ld >= V, saved_ld % 8 == 0, at least T * saved_ld bf16 elements.
 * ---------------------------------------------------------------------------------------- */
int64_t prl_lmhead_saved_logits_ld(int64_t V);
int prl_lmhead_fwd_save(const void* hidden, const void* weight, const int64_t* ids, int64_t T, int32_t H, int64_t V,
                        int64_t v_offset, int64_t V_total, float inv_temp, int32_t shift, float* logp,
                        float* entropy, float* lse, float* partials, void* saved_logits, int64_t saved_ld,
                        void* workspace, size_t workspace_bytes, void* stream);
int prl_lmhead_bwd_saved(const void* hidden, const void* weight, const int64_t* ids, const float* lse,
                         const float* dlogp, int64_t T, int32_t H, int64_t V, int64_t v_offset, float inv_temp,
                         int32_t shift, void* saved_logits, int64_t saved_ld, void* d_hidden,
                         int32_t d_hidden_f32, void* d_weight, int32_t d_weight_dtype, int32_t accumulate_dw,
                         int32_t phase_mask, void* workspace, size_t workspace_bytes, void* stream);
size_t prl_lmhead_bwd_saved_workspace_bytes(int64_t T, int32_t H, int64_t V);

/* ------------------------------------------------------------------------------------------
 * Masked positions (prompt / environment tokens).  The reference multiplies their loss by the mask
 * (src/prime_rl/trainer/rl/loss.py:104; mask built by src/prime_rl/orchestrator/batch.py:41-44), so their dlogits rows
 * are zero: the LM-head GEMMs only need the ACTIVE positions.  These entry points replace host-side index tensors
 * (nonzero / index_select / index_copy_) by device kernels:
 *   prl_partition_positions: stable partition of positions [shift, T) of a packed row by mask (uint8, torch.bool storage):
 *     order[j] = position in slot j (active ones first, then the others, each in order), j < T - shift;
 *     slot[t] = slot of position t (-1 for t < shift: shift_logits' zero row has no source row); *n_active (device int32).
 *   prl_gather_rows_bf16:  dst[j, :H] = src[order[j] - shift, :H] for j < n (bf16 rows, H % 8 == 0, 16-byte aligned)
 *     and ids_out[j] = ids[order[j]] (ids / ids_out may be NULL).
 *   prl_gather_tokens_f32: x_out[j] = x[order[j]], j < n, for up to three fp32 vectors (NULL pairs are skipped).
 *   prl_scatter_tokens_f32: x_out[t] = x[slot[t]] if 0 <= slot[t] < n_valid, x_first if slot[t] < 0, else 0; t < T.
 *   prl_scatter_rows_bf16: dst[r, :H] = src[slot[r + shift], :H] if that slot is in [0, n_valid), else zeros; r < T
 *     (the gradient w.r.t. hidden back in source-row order; masked rows and the last `shift` rows get exact zeros).
 * ---------------------------------------------------------------------------------------- */
int prl_partition_positions(const uint8_t* mask, int64_t T, int32_t shift, int32_t* order, int32_t* slot, int32_t* n_active,
                            void* stream);
int prl_gather_rows_bf16(const void* src, const int64_t* ids, const int32_t* order, int64_t n, int32_t shift, int32_t H, void* dst,
                         int64_t* ids_out, void* stream);
int prl_gather_tokens_f32(const int32_t* order, int64_t n, const float* a, float* a_out, const float* b, float* b_out, const float* c,
                          float* c_out, void* stream);
int prl_scatter_tokens_f32(const int32_t* slot, int64_t T, int64_t n_valid, const float* a, float* a_out, float a_first, const float* b,
                           float* b_out, float b_first, const float* c, float* c_out, float c_first, void* stream);
int prl_scatter_rows_bf16(const void* src, const int32_t* slot, int64_t T, int64_t n_valid, int32_t shift, int32_t H, void* dst,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device logging statistics ("next" row N1).  Replaces the per-micro-batch boolean-index + blocking .to("cpu")
 * of src/prime_rl/trainer/rl/train.py:313-333 and the padded all-gather of every masked token in
 * Tensors.compute_stats (src/prime_rl/trainer/utils.py:228-256) by five running moments per logged key:
 *   acc[0..4] (fp64, device) = (count, sum, sum of squares, min, max) over the elements with mask != 0
 * (mask == NULL: all elements) of f(x), f = identity (transform 0) or exp (transform 1, e.g. probs = exp(logprobs)).
 * acc is UPDATED in place (initialise it to {0, 0, 0, +inf, -inf}); ranks then all-reduce 5 scalars per key.
 * ---------------------------------------------------------------------------------------- */
size_t prl_masked_moments_workspace_bytes(void);
int prl_masked_moments_f32(const float* x, const uint8_t* mask, int64_t n, int32_t transform, double* acc, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Exact median of the logged values without gathering them (N1, `torch.median` at trainer/utils.py:248 = the lower
 * median, sorted[(n-1)/2]).  prl_masked_keys_u32 appends an order-preserving uint32 key of f(x[i]) for every unmasked
 * element to keys[*counter ...] (device counter, advanced atomically; elements beyond `capacity` are dropped, so size
 * the buffer for the number of elements offered).  A 3-pass radix select (bits 31..21, 20..10, 9..0) then finds the
 * k-th smallest key: prl_radix_select_hist adds to hist[2048] (uint64, device) the histogram of bits
 * [shift, shift + bits) over the keys that match state = {prefix, prefix_mask, k, -} (int64[4], device);
 * ranks all-reduce(SUM) the histogram; prl_radix_select_pick locates the bin of the k-th element, extends
 * prefix / prefix_mask, reduces k, and clears hist.  After the last pass state[0] is the key of the median
 * (key -> float: key & 0x80000000 ? key ^ 0x80000000 : ~key).  No host synchronisation anywhere. */
int prl_masked_keys_u32(const float* x, const uint8_t* mask, int64_t n, int32_t transform, uint32_t* keys,
                        int64_t capacity, unsigned long long* counter, void* stream);
int prl_radix_select_hist(const uint32_t* keys, const unsigned long long* counter, int64_t capacity,
                          const int64_t* state, int32_t shift, int32_t bits, unsigned long long* hist, void* stream);
int prl_radix_select_pick(unsigned long long* hist, int64_t* state, int32_t shift, int32_t bits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Packed micro-batch rows from rollouts ("next" row N4).  Replaces the per-token tensor building of prepare_sample
 * (src/prime_rl/orchestrator/batch.py:21-62: cat(prompt, completion), arange position ids, advantage.repeat(len),
 * cat(zeros(len(prompt)), completion_logprobs)) and the concatenation of prepare_micro_batch_packing (:105-120) for ALL
 * rollouts of a step in one launch.  Rollout i (source order) has tokens[src_tok_off[i] .. src_tok_off[i+1]) =
 * prompt + completion (int64), mask[...] (uint8, same indexing), completion log-probs lp[src_lp_off[i] ..), prompt_len[i],
 * advantage[i]; it is written at dst_off[i] of the packed stream (the bin order decided by the host's first-fit-decreasing
 * pass, batch.py:76-102).  Outputs are flat [total] arrays; a micro-batch is a contiguous slice of them.
 * ---------------------------------------------------------------------------------------- */
int prl_pack_rollouts(const int64_t* tokens, const uint8_t* mask, const float* completion_logprobs, const int64_t* src_tok_off,
                      const int64_t* src_lp_off, const int32_t* prompt_len, const float* advantage, const int64_t* dst_off, int64_t n,
                      int64_t* input_ids, int64_t* position_ids, float* advantages, float* logprobs, uint8_t* loss_mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Plain bf16 GEMM on the same tcgen05 engine, exposed for tests and bring-up:
 *   C[M,N] (fp32) = A x B with per-operand major-ness:
 *   a_mn_major == 0: A is [M,K] row-major (K contiguous); != 0: A is stored [K,M] (M contiguous)
 *   b_mn_major == 0: B is [N,K] row-major (K contiguous); != 0: B is stored [K,N] (N contiguous)
 * ---------------------------------------------------------------------------------------- */
int prl_gemm_bf16_f32(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int32_t a_mn_major,
                      int32_t b_mn_major, int32_t accumulate, void* stream);

/* Tuning knobs.  Every knob has a compiled-in default; the environment variable of the same name is
 * read once, when the library first needs a knob; prl_set_knob changes it afterwards (process-wide; meant for tests and tuning
 * scripts, not for production use).  Names: PRL_VCHUNK, PRL_FWD_STAGES, PRL_DZ_STAGES, PRL_GROUP_M, PRL_SWEEP_GROUP_M,
 * PRL_TILE_STAGES, PRL_K_SPLITS, PRL_L2_PROMO, PRL_EPI_TMA, PRL_HINT_<FWD|FWDS|DZ|DH|DW|GEMM>_<A|B|C>, and -- before the
 * first compute call only -- PRL_CTA_GROUP, PRL_MSUB.  Profiling aids (PRL_DEBUG_*) exist only in -DPRL_DEBUG_KNOBS builds. */
int prl_set_knob(const char* name, int64_t value);
int prl_get_knob(const char* name, int64_t* value);

/* Test hook (host only, no GPU): the work decomposition the GEMM engine would use for an M x N x K problem.
 * kind 0 = forward sweep, 1 = dlogits sweep, 2 / 3 = store tiles of the recompute / saved-logits backward (incl. the
 * K-split decision).  items[6 * i ..] = (m_tile, n_begin, n_count, k_begin, k_count, split) for item i (up to `capacity`
 * items; may be NULL); dims[8] = (m_tiles, n_tiles, k_blocks, num_items, workers, k_splits, n_splits, rows per m-tile). */
int prl_debug_schedule(int32_t kind, int64_t M, int64_t N, int64_t K, int32_t* items, int64_t capacity, int32_t* dims);

/* Test hook (host only): the tile order of the fused reduce-scatter d_hidden GEMM (prl_lmhead_bwd_scatter) for rank `self_index`
 * of `n_peers`, M = n_peers * rows_per_peer output rows.  items as in prl_debug_schedule; dims[8] = (m_tiles, n_tiles, k_blocks,
 * num_items, workers, m_rot = m-tile rotation, rows_per_peer, rows per m-tile). */
int prl_debug_schedule_scatter(int64_t M, int64_t N, int64_t K, int32_t n_peers, int32_t self_index, int32_t* items,
                               int64_t capacity, int32_t* dims);

/* number of kernel launches issued by this library on this host thread since the last reset
 * (bench.py's `gpu_launches`) */
int64_t prl_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* PRL_B200_H_ */
