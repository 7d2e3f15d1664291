"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Token-sharded mode (default; = FSDP ranks): micro-batches are dealt to ranks in contiguous chunks
(src/prime_rl/orchestrator/batch.py:167-173) and the loss is normalised by the rank-LOCAL scale
(src/prime_rl/trainer/rl/train.py:263-267), so the path needs no collective of its own.

Vocab-sharded LM head (BASELINE.json config 4; not in the reference): each rank owns rows
[v0, v1) of W_lm (FSDP2's dim-0 chunking) and sees all T tokens; ranks exchange only the per-token
softmax partial (m, s, u, z_target) = 16 bytes/token, then d_hidden partial sums are all-reduced.
"""

from __future__ import annotations

from typing import Any, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .peer import PeerBuffer, ShardedGradAccumulator  # noqa: F401  (re-exported)
from .lm_head import (lm_head_backward_raw, lm_head_backward_saved_raw, lm_head_forward_raw, lm_head_forward_save_raw,
                      saved_logits_policy, _warn_fallback)
from .rows import first_row_values


def rows_first_lse(V: int) -> float:
    """lse of a position whose logits are all zero (zero hidden row): ln V"""
    return first_row_values(V)[2]


# --------------------------------------------------------------------------- token sharding (host logic)
def shard_micro_batches(micro_batches: Sequence[Any], rank: int, world_size: int) -> list[Any]:
    """Rank r's contiguous chunk of len(micro_batches) // world_size items (batch.py:167-173: the
    orchestrator pads the list to a multiple of the worker count first, batch.py:153-161)."""
    n = len(micro_batches)
    if n % world_size != 0:
        raise ValueError("number of micro-batches must be a multiple of the number of ranks (batch.py:153)")
    per = n // world_size
    return list(micro_batches[rank * per : (rank + 1) * per])


def local_loss_scale(micro_batches: Sequence[Any], ratio_type: str) -> int:
    """train.py:260-267: token mode = number of unmasked tokens in THIS rank's micro-batches; sequence mode
    = micro_batch_size (rows of input_ids) x number of micro-batches."""
    if ratio_type == "token":
        return int(sum(int(mb["loss_mask"].sum().item()) for mb in micro_batches))
    if ratio_type == "sequence":
        return int(micro_batches[0]["input_ids"].shape[0] * len(micro_batches)) if len(micro_batches) else 0
    raise ValueError(ratio_type)


def vocab_shard_bounds(vocab: int, world_size: int, rank: int) -> tuple[int, int]:
    """Rows [v0, v1) of W_lm owned by `rank` under FSDP2 dim-0 chunking (torch.chunk semantics: chunks of
    ceil(V / world) rows, trailing ranks may get fewer or none).  151936 / 8 = 18992 rows per rank."""
    per = -(-vocab // world_size)
    v0 = min(rank * per, vocab)
    return v0, min(v0 + per, vocab)


def exchange_partials(partials: torch.Tensor, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """all-gather of the per-token partials: [T, 4] on every rank -> [world, T, 4] (rank-major)."""
    world = dist.get_world_size(group)
    T = partials.shape[0]
    out = torch.empty((world * T,) + tuple(partials.shape[1:]), dtype=partials.dtype, device=partials.device)
    dist.all_gather_into_tensor(out, partials.contiguous(), group=group)  # concatenation along dim 0, rank-major
    return out.view((world,) + tuple(partials.shape))


# --------------------------------------------------------------------------- vocab-sharded fused LM head
class _VocabParallelLMHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight_shard, ids, inv_temp, v_offset, v_total, group, shift, round_bf16, save_logits=None):
        lib = _lib.load()
        h2 = hidden.detach().reshape(-1, hidden.shape[-1]).contiguous()
        w2 = weight_shard.detach().contiguous()
        i2 = ids.detach().reshape(-1).contiguous().to(torch.int64)
        T = h2.shape[0]
        dev = h2.device
        ctx.saved_logits = None  # this shard's bf16 logits [T, V_local], consumed by the backward
        save = (round_bf16 and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
                and saved_logits_policy(T, w2.shape[0], save_logits))
        if w2.shape[0] > 0 and save:
            try:
                _, _, _, parts, ctx.saved_logits = lm_head_forward_save_raw(h2, w2, i2, inv_temp, shift, v_offset=v_offset,
                                                                            v_total=v_total, want_partials=True)
            except torch.OutOfMemoryError:
                _warn_fallback(f"no room for the {T} x {w2.shape[0]} bf16 logits block of this vocab shard")
                save = False
        if w2.shape[0] > 0 and not save:
            _, _, _, parts = lm_head_forward_raw(h2, w2, i2, inv_temp, shift, round_bf16, v_offset=v_offset, v_total=v_total,
                                                 want_partials=True)
        elif w2.shape[0] == 0:  # a trailing rank with an empty shard contributes the neutral element
            parts = torch.zeros(T, 4, dtype=torch.float32, device=dev)
            parts[:, 0] = float("-inf")
        gathered = exchange_partials(parts, group)
        logp = torch.empty(T, dtype=torch.float32, device=dev)
        ent = torch.empty(T, dtype=torch.float32, device=dev)
        lse = torch.empty(T, dtype=torch.float32, device=dev)
        rc = lib.prl_lmhead_merge_partials(_lib.ptr(gathered), gathered.shape[0], T, _lib.ptr(logp), _lib.ptr(ent), _lib.ptr(lse),
                                           _lib.stream())
        _lib.check(rc, "prl_lmhead_merge_partials")
        ctx.save_for_backward(h2, w2, i2, lse)
        ctx.cfg = (inv_temp, v_offset, group, shift, round_bf16, hidden.shape)
        lead = hidden.shape[:-1]
        logp, ent = logp.reshape(lead), ent.reshape(lead)
        ctx.mark_non_differentiable(ent)
        return logp, ent

    @staticmethod
    def backward(ctx, g_logp, _g_ent):
        h2, w2, i2, lse = ctx.saved_tensors
        inv_temp, v_offset, group, shift, round_bf16, hshape = ctx.cfg
        g = g_logp.detach().reshape(-1).contiguous().float()
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dh = dw = None
        saved, ctx.saved_logits = ctx.saved_logits, None
        if w2.shape[0] > 0 and saved is not None:
            dh, dw = lm_head_backward_saved_raw(h2, w2, i2, lse, g, inv_temp, saved, shift, v_offset=v_offset, want_dhidden=need_h,
                                                want_dweight=need_w, dhidden_f32=True)
        elif w2.shape[0] > 0:
            dh, dw = lm_head_backward_raw(h2, w2, i2, lse, g, inv_temp, shift, round_bf16, v_offset=v_offset, want_dhidden=need_h,
                                          want_dweight=need_w, dhidden_f32=True)
        elif need_h:
            dh = torch.zeros(h2.shape, dtype=torch.float32, device=h2.device)
        if need_h:
            dist.all_reduce(dh, group=group)  # partial sums over vocab shards
            dh = dh.to(torch.bfloat16).reshape(hshape)
        if need_w and dw is None:
            dw = torch.zeros_like(w2)
        return dh, dw, None, None, None, None, None, None, None, None


def vocab_parallel_lm_head_logprobs_entropy(
    hidden: torch.Tensor,
    weight_shard: torch.Tensor,
    input_ids: torch.Tensor,
    temperature: float,
    vocab_size: int,
    group: dist.ProcessGroup | None = None,
    *,
    shift: bool = True,
    round_bf16: bool = True,
    save_logits: bool | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """Vocab-sharded variant of `fused_lm_head_logprobs_entropy`: `weight_shard` holds this rank's rows
    `vocab_shard_bounds(vocab_size, world, rank)` of W_lm; `hidden` / `input_ids` are the SAME on every rank
    of `group` (all-gathered by the caller).  Returns (logprobs, entropy) identical on every rank; backward
    gives the full d_hidden (all-reduced) and the local shard of d_W_lm.  `save_logits`: as in
    `fused_lm_head_logprobs_entropy`, applied to this rank's `[T, V_local]` block."""
    _lib.require_cuda(hidden, weight_shard, input_ids)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    v0, v1 = vocab_shard_bounds(vocab_size, world, rank)
    if weight_shard.shape[0] != v1 - v0:
        raise ValueError(f"rank {rank} expects {v1 - v0} vocab rows, got {weight_shard.shape[0]}")
    return _VocabParallelLMHead.apply(hidden, weight_shard, input_ids, 1.0 / float(temperature), v0, vocab_size, group,
                                      bool(shift), bool(round_bf16), save_logits)


# --------------------------------------------------------------------------- config 4: rounds of one micro-batch per rank
def partials_to_owners(parts: torch.Tensor, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """parts [world, T, 4]: this rank's vocab-shard partial (m, s, u, z_target) for the T tokens of every rank (token-owner
    major).  Returns [world, T, 4]: the partials of THIS rank's tokens from every vocab shard (shard major) -- an all-to-all of
    16 bytes per token and shard, the only forward exchange of the vocab-sharded head."""
    mine = torch.empty_like(parts)
    dist.all_to_all_single(mine.view(-1), parts.contiguous().view(-1), group=group)
    return mine


def gather_token_scalars(lse_local: torch.Tensor, dlogp_local: torch.Tensor, group: dist.ProcessGroup | None = None):
    """(lse, d loss / d logp) of every rank's tokens, 8 bytes per token: what the backward sweep over a vocab shard needs for
    tokens it does not own.  Position 0 of every packed row is shift_logits' constant row: its dlogp is forced to 0 so that no
    gradient reaches the previous rank's last hidden row in the concatenated sweep.  Returns (lse_all, dlogp_all) [world * T]."""
    T = lse_local.numel()
    world = dist.get_world_size(group)
    pair = torch.stack([lse_local.float().reshape(-1), dlogp_local.float().reshape(-1)], dim=1).contiguous()  # [T, 2]
    if T > 0:
        pair[0, 1] = 0.0
    pair_all = torch.empty(world * T, 2, dtype=torch.float32, device=pair.device)
    dist.all_gather_into_tensor(pair_all, pair, group=group)
    return pair_all[:, 0].contiguous(), pair_all[:, 1].contiguous()


class VocabShardedLMHead:
    """BASELINE.json config 4 as the trainer would run it: the LM head stays sharded over the vocabulary exactly as FSDP2 stores
    it (dim-0 chunks of `lm_head.weight`, src/prime_rl/trainer/model.py:100-135: no un-shard, no full d_W per rank), every rank
    brings ITS OWN packed micro-batch of T tokens, and one ROUND processes the `world` micro-batches together:

        all-gather hidden [world, T, H] (+ ids)              in-path, NCCL; can be issued ahead (`gather_inputs`)
        local GEMM sweep over all world*T tokens x V/world columns  -> per-token partial (m, s, u, z_target)
        all-to-all of the partials, 16 B per token and shard  -> every rank merges the partials of ITS tokens
        packed loss on the local tokens                      -> d loss / d logp (local)
        all-gather (lse, dlogp), 8 B per token
        backward sweep: d_W shard complete locally (fp32 accumulation, already in FSDP's shard layout);
                        d_hidden partial sums [world*T, H] are REDUCE-SCATTERED to the token owners: either by NCCL after
                        the GEMM (fused_rs=False) or inside the d_hidden GEMM's epilogue, tile by tile, as TMA reduce-adds into
                        the owner's symmetric-memory accumulator over NVLink (fused_rs=True).

    The exchanges need the same row length on every rank.  Packed rows are NOT padded by the orchestrator (batch.py:115-116:
    T = sum of the sample lengths <= seq_len), so `gather_inputs(..., pad_to=seq_len)` pads this rank's row to the common length
    with zero hidden rows whose d loss / d logp is 0: they add exact zeros to d_W and d_hidden, and `forward` / `backward` /
    `micro_step` return the local T tokens only.  Without `pad_to` every rank must bring the same T."""

    def __init__(self, weight_shard: torch.Tensor, vocab_size: int, group: dist.ProcessGroup | None = None, *,
                 main_grad: torch.Tensor | None = None, fused_rs: bool = False, round_bf16: bool = True):
        _lib.require_cuda(weight_shard)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.V = int(vocab_size)
        self.v0, self.v1 = vocab_shard_bounds(self.V, self.world, self.rank)
        if weight_shard.shape[0] != self.v1 - self.v0 or weight_shard.dtype != torch.bfloat16 or not weight_shard.is_contiguous():
            raise ValueError(f"rank {self.rank} expects a contiguous bf16 [{self.v1 - self.v0}, H] shard of W_lm")
        if self.v1 - self.v0 == 0:
            raise ValueError("empty vocab shard (more ranks than vocab rows / chunk)")
        self.W = weight_shard
        self.H = weight_shard.shape[1]
        self.main_grad = main_grad
        self.round_bf16 = round_bf16
        self.fused_rs = bool(fused_rs)
        self.comm = torch.cuda.Stream(device=weight_shard.device)
        self._sym = None      # symmetric-memory accumulator of the fused reduce-scatter
        self._zero_row = None
        self._ws = {}
        self.timings = {}     # CUDA events of the last round's exchanges

    # -- inputs: all-gather of hidden and ids (can run ahead of the previous round's compute on the comm stream)
    def gather_inputs(self, hidden_local: torch.Tensor, ids_local: torch.Tensor, pad_to: int | None = None):
        T_local, H = hidden_local.shape[-2], hidden_local.shape[-1]
        dev = hidden_local.device
        h = hidden_local.detach().reshape(T_local, H).contiguous()
        i = ids_local.detach().reshape(T_local).contiguous().to(torch.int64)
        T = T_local if pad_to is None else int(pad_to)
        if T < T_local:
            raise ValueError(f"pad_to = {T} is shorter than this rank's row of {T_local} tokens")
        if T > T_local:  # ragged rows: zero hidden rows / token id 0 up to the common length (their dlogp is 0 in the backward)
            h_pad = torch.zeros(T, H, dtype=h.dtype, device=dev)
            h_pad[:T_local] = h
            i_pad = torch.zeros(T, dtype=torch.int64, device=dev)
            i_pad[:T_local] = i
            h, i = h_pad, i_pad
        h_all = torch.empty(self.world * T, H, dtype=torch.bfloat16, device=dev)
        i_all = torch.empty(self.world * T, dtype=torch.int64, device=dev)
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            dist.all_gather_into_tensor(h_all, h, group=self.group)
            dist.all_gather_into_tensor(i_all, i, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        for t in (h, i, h_all, i_all):
            t.record_stream(self.comm)
        return dict(h_all=h_all, i_all=i_all, T=T, T_local=T_local, ready=ev)

    def _workspace(self, key, nbytes, dev):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return ws

    def forward(self, gathered, temperature: float):
        """-> (logp, entropy, lse) of the LOCAL tokens [T] (merged over all vocab shards)"""
        lib = _lib.load()
        T, W = gathered["T"], self.world
        h_all, i_all = gathered["h_all"], gathered["i_all"]
        dev = h_all.device
        torch.cuda.current_stream().wait_event(gathered["ready"])
        inv = 1.0 / float(temperature)
        ws = self._workspace("fwd", lib.prl_lmhead_fwd_workspace_bytes(W * T, self.H, self.v1 - self.v0), dev)
        # one sweep over the concatenated rows; the shift is per packed row, so the first position of every rank's row is
        # shift_logits' all-zero row (rl/loss.py:55-56), not the previous rank's last token: closed-form partial (0, V_shard, 0, 0)
        _, _, _, parts = lm_head_forward_raw(h_all, self.W, i_all, inv, True, self.round_bf16, v_offset=self.v0, v_total=self.V,
                                             want_partials=True, workspace=ws)
        parts = parts.view(W, T, 4)
        if W > 1:
            if self._zero_row is None:
                self._zero_row = torch.tensor([0.0, float(self.v1 - self.v0), 0.0, 0.0], device=dev)
            parts[1:, 0] = self._zero_row
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mine = partials_to_owners(parts, self.group)  # [shard, T, 4]: the partials of MY tokens from every vocab shard
        e1.record()
        self.timings["partials_all_to_all"] = (e0, e1)
        logp = torch.empty(T, dtype=torch.float32, device=dev)
        ent = torch.empty(T, dtype=torch.float32, device=dev)
        lse = torch.empty(T, dtype=torch.float32, device=dev)
        rc = lib.prl_lmhead_merge_partials(_lib.ptr(mine), W, T, _lib.ptr(logp), _lib.ptr(ent), _lib.ptr(lse), _lib.stream())
        _lib.check(rc, "prl_lmhead_merge_partials")
        n = gathered.get("T_local", T)
        if n < T:
            gathered["lse_row"] = lse  # the padded positions' lse (ln V: zero logits) is what the backward sweep reads for them
            return logp[:n], ent[:n], lse[:n]
        return logp, ent, lse

    def backward(self, gathered, temperature: float, lse_local: torch.Tensor, dlogp_local: torch.Tensor, vchunk: int = 0):
        """-> d_hidden of the LOCAL tokens [T, H] bf16; the d_W shard is accumulated into `main_grad` (fp32 [V_local, H])"""
        T, W = gathered["T"], self.world
        h_all, i_all = gathered["h_all"], gathered["i_all"]
        dev = h_all.device
        inv = 1.0 / float(temperature)
        n = gathered.get("T_local", T)
        if lse_local.numel() != n or dlogp_local.numel() != n:
            raise ValueError(f"lse / dlogp must cover this rank's {n} tokens")
        if n < T:  # ragged row: padded positions carry dlogp = 0 (exact-zero dlogits rows) and the forward's own lse
            g_row = torch.zeros(T, dtype=torch.float32, device=dev)
            g_row[:n] = dlogp_local.reshape(-1)
            lse_row = gathered.get("lse_row")
            if lse_row is None:
                lse_row = torch.full((T,), rows_first_lse(self.V), dtype=torch.float32, device=dev)
                lse_row[:n] = lse_local.reshape(-1)
            lse_local, dlogp_local = lse_row, g_row
        if self.fused_rs:
            # the owner's accumulator must be zero before any peer's d_hidden GEMM adds into it: zeroed HERE, i.e. before this
            # rank enters the all-gather below, which no rank can leave before every rank has entered it
            self._peer_accumulator(T, dev).zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lse_all, g_all = gather_token_scalars(lse_local, dlogp_local, self.group)
        e1.record()
        self.timings["lse_dlogp_all_gather"] = (e0, e1)
        if vchunk <= 0:
            vchunk = (self.v1 - self.v0 + 255) // 256 * 256  # the whole shard in one chunk: one dlogits block, one d_hidden pass
        lib = _lib.load()
        ws = self._workspace("bwd", lib.prl_lmhead_bwd_workspace_bytes(W * T, self.H, self.v1 - self.v0, vchunk), dev)
        if self.main_grad is None:
            self.main_grad = torch.zeros(self.v1 - self.v0, self.H, dtype=torch.float32, device=dev)
        if self.fused_rs:
            return self._backward_fused(gathered, inv, lse_all, g_all, vchunk, ws)[:n]
        dh_part, _ = lm_head_backward_raw(h_all, self.W, i_all, lse_all, g_all, inv, True, self.round_bf16, v_offset=self.v0,
                                          dweight=self.main_grad, dweight_accumulate=True, vchunk=vchunk, dhidden_f32=True, workspace=ws)
        dh_local = torch.empty(T, self.H, dtype=torch.float32, device=dev)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        dist.reduce_scatter_tensor(dh_local, dh_part, group=self.group)  # partial sums over vocab shards -> token owners
        r1.record()
        self.timings["d_hidden_reduce_scatter"] = (r0, r1)
        return dh_local[:n].to(torch.bfloat16)

    # -- fused reduce-scatter: peer-mapped fp32 accumulators, one per rank, written by every rank's d_hidden epilogue
    def _peer_accumulator(self, T: int, dev) -> torch.Tensor:
        if self._sym is not None and self._sym.rows == T:
            return self._sym.local
        if self._sym is not None:
            self.close()
        if T % 32 != 0:
            raise ValueError("fused reduce-scatter needs T % 32 == 0")
        try:
            self._sym = PeerBuffer(T, self.H, dev, self.group)
        except RuntimeError as e:
            raise RuntimeError(f"fused reduce-scatter unavailable ({e}); use fused_rs=False") from e
        return self._sym.local

    def _backward_fused(self, gathered, inv, lse_all, g_all, vchunk, ws):
        lib = _lib.load()
        T, W = gathered["T"], self.world
        h_all, i_all = gathered["h_all"], gathered["i_all"]
        sym = self._sym
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dt = _lib.PRL_DT_F32 if self.main_grad.dtype == torch.float32 else _lib.PRL_DT_BF16
        rc = lib.prl_lmhead_bwd_scatter(_lib.ptr(h_all), _lib.ptr(self.W), _lib.ptr(i_all), _lib.ptr(lse_all), _lib.ptr(g_all), W * T, self.H,
                                        self.v1 - self.v0, self.v0, float(inv), 1, 1 if self.round_bf16 else 0, sym.ptrs, W, T, self.rank,
                                        _lib.ptr(self.main_grad), dt, 1, vchunk, 0, _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "prl_lmhead_bwd_scatter")
        # all ranks' epilogues must have landed before the accumulator is read: a 4-byte all-reduce is the cross-rank barrier
        r0.record()
        sym.barrier()
        r1.record()
        self.timings["d_hidden_reduce_scatter"] = (r0, r1)  # fused: what is left of it outside the GEMM = the barrier
        return sym.local.to(torch.bfloat16)

    def close(self):
        """unmap the peers' accumulators (fused reduce-scatter)"""
        if self._sym is not None:
            self._sym.close()
            self._sym = None

    def micro_step(self, hidden_local, micro_batch, loss_config, loss_scale, gathered=None, pad_to=None):
        """One round from this rank's point of view (train.py:285-310 with the head vocab-sharded): returns loss, logprobs,
        entropy, d_hidden of the LOCAL micro-batch; the d_W shard is accumulated into `main_grad`.  `pad_to` (= seq_len): the
        common row length when the ranks' packed rows differ in length (see the class docstring)."""
        from .loss import compute_loss
        from .segments import get_response_lengths

        ids = micro_batch["input_ids"]
        if gathered is None:
            gathered = self.gather_inputs(hidden_local, ids, pad_to)
        lengths = get_response_lengths(micro_batch["position_ids"])
        tau = micro_batch["temperature"]
        logp, ent, lse = self.forward(gathered, tau)
        lp = logp.detach().requires_grad_(True)
        loss, tensors = compute_loss(lp.split(lengths), micro_batch["logprobs"].reshape(-1).split(lengths),
                                     micro_batch["advantages"].reshape(-1).split(lengths), micro_batch["loss_mask"].reshape(-1).split(lengths),
                                     loss_config, loss_scale)
        loss.backward()
        dh = self.backward(gathered, tau, lse, lp.grad)
        out = {"loss": loss.detach(), "logprobs": logp, "entropy": ent, "d_hidden": dh}
        out.update(tensors)
        return out

    def exchange_ms(self) -> dict:
        """duration (ms) of the last round's exchange steps, from CUDA events on the launching stream (synchronises)"""
        torch.cuda.synchronize()
        return {k: a.elapsed_time(b) for k, (a, b) in self.timings.items()}
