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
from .lm_head import (lm_head_backward_raw, lm_head_backward_saved_raw, lm_head_forward_raw, lm_head_forward_save_raw,
                      saved_logits_policy)


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
