"""Orchestrator-side packing on the device ("next" row N4 of SURVEY.md §8f): a drop-in for
`prepare_batch` (src/prime_rl/orchestrator/batch.py:123-175).

The reference builds five small tensors per rollout in Python (`prepare_sample`, :21-62), deep-copies every rollout, packs
with first-fit-decreasing (:76-102) and concatenates per bin (:105-120).  Here the bin assignment stays on the host (it only
looks at lengths) and ONE kernel writes every per-token tensor of every micro-batch of the step (`prl_pack_rollouts`); a
micro-batch is a contiguous slice of the flat outputs.
"""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from . import _lib


def first_fit_decreasing(lengths: Sequence[int], max_seq_len: int) -> list[list[int]]:
    """`packed_samples_into_micro_bs` (batch.py:76-102) on lengths only: stable sort by length, longest first, each sample
    into the first bin it fits.  Returns the bins as lists of sample indices (bins in creation order)."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i], reverse=True)
    bins: list[list[int]] = []
    fill: list[int] = []
    for i in order:
        for b in range(len(bins)):
            if fill[b] + lengths[i] <= max_seq_len:
                bins[b].append(i)
                fill[b] += lengths[i]
                break
        else:
            bins.append([i])
            fill.append(lengths[i])
    return bins


def prepare_batch(
    rollouts: Sequence[Any],
    temperature: float,
    tokenizer: Any,
    batch_size: int,
    micro_batch_size: int,
    seq_len: int,
    num_train_workers: int,
    *,
    device_output: bool = False,
) -> list[list[dict]]:
    """Same arguments and result as the reference: one list of `MicroBatch` dicts per train worker, each `[1, T]` packed
    row un-padded, padding micro-batches (copies of the first one with zero advantages and an all-False mask) appended so
    every worker gets the same count (batch.py:153-161), contiguous chunks per worker (:167-173).  `rollouts` are objects
    with the fields of `Rollout` (orchestrator/buffer.py:19-30).  Tensors are CPU tensors like the reference's (they are
    `torch.save`d for the trainer) unless `device_output=True`."""
    del tokenizer, batch_size  # unused by the reference as well
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    max_seq_len = seq_len * micro_batch_size
    n = len(rollouts)
    if n == 0:
        raise IndexError("prepare_batch needs at least one rollout")  # the reference fails on micro_batches[0]
    p_len = np.asarray([len(r.prompt_tokens) for r in rollouts], dtype=np.int64)
    c_len = np.asarray([len(r.completion_tokens) for r in rollouts], dtype=np.int64)
    lens = p_len + c_len
    for ln in lens:
        if ln > max_seq_len:
            raise ValueError(f"Number of tokens {ln} is greater than sequence length {max_seq_len}. This should not happen.")
    bins = first_fit_decreasing(lens.tolist(), max_seq_len)

    # flat source arrays (source order) and the destination offset of every rollout in the packed stream (bin order)
    tokens = np.concatenate([np.asarray(list(r.prompt_tokens) + list(r.completion_tokens), dtype=np.int64) for r in rollouts])
    mask = np.concatenate([np.asarray(list(r.prompt_mask) + list(r.completion_mask), dtype=np.int64) for r in rollouts]).astype(np.uint8)
    lp = np.concatenate([np.asarray(r.completion_logprobs, dtype=np.float32) for r in rollouts] + [np.zeros(1, np.float32)])
    if any(len(r.completion_logprobs) != c for r, c in zip(rollouts, c_len)):
        raise AssertionError("completion_logprobs and completion_tokens disagree")  # batch.py:53-55
    src_tok_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    src_lp_off = np.concatenate([[0], np.cumsum(c_len)]).astype(np.int64)
    dst_off = np.zeros(n, dtype=np.int64)
    mb_bounds = [0]
    o = 0
    for b in bins:
        for i in b:
            dst_off[i] = o
            o += int(lens[i])
        mb_bounds.append(o)
    total = o
    adv = np.asarray([r.advantage for r in rollouts], dtype=np.float32)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().to(dev, non_blocking=True)

    d_tokens, d_mask, d_lp = up(tokens), up(mask), up(lp)
    d_sto, d_slo, d_pl, d_adv, d_dst = up(src_tok_off), up(src_lp_off), up(p_len.astype(np.int32)), up(adv), up(dst_off)
    input_ids = torch.empty(total, dtype=torch.int64, device=dev)
    position_ids = torch.empty(total, dtype=torch.int64, device=dev)
    advantages = torch.empty(total, dtype=torch.float32, device=dev)
    logprobs = torch.empty(total, dtype=torch.float32, device=dev)
    loss_mask = torch.empty(total, dtype=torch.bool, device=dev)
    rc = lib.prl_pack_rollouts(_lib.ptr(d_tokens), _lib.ptr(d_mask), _lib.ptr(d_lp), _lib.ptr(d_sto), _lib.ptr(d_slo), _lib.ptr(d_pl),
                               _lib.ptr(d_adv), _lib.ptr(d_dst), n, _lib.ptr(input_ids), _lib.ptr(position_ids), _lib.ptr(advantages),
                               _lib.ptr(logprobs), _lib.ptr(loss_mask.view(torch.uint8)), _lib.stream())
    _lib.check(rc, "prl_pack_rollouts")
    flat = dict(input_ids=input_ids, advantages=advantages, loss_mask=loss_mask, position_ids=position_ids, logprobs=logprobs)
    if not device_output:
        flat = {k: v.cpu() for k, v in flat.items()}

    micro_batches = []
    for k in range(len(bins)):
        a, b = mb_bounds[k], mb_bounds[k + 1]
        mb = {key: v[a:b].unsqueeze(0) for key, v in flat.items()}
        mb["temperature"] = temperature
        micro_batches.append(mb)

    num_padding_batch = num_train_workers - len(micro_batches) % num_train_workers  # batch.py:153 (never 0)
    if num_train_workers > 1 and num_padding_batch > 0:
        first = micro_batches[0]
        padded = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in first.items()}
        padded["advantages"] = torch.zeros_like(padded["advantages"])
        padded["loss_mask"] = torch.zeros_like(padded["loss_mask"], dtype=torch.bool)
        micro_batches.extend([padded for _ in range(num_padding_batch)])
    assert len(micro_batches) % num_train_workers == 0, "Number of micro batches is not divisible by number of data ranks"
    per = len(micro_batches) // num_train_workers
    return [micro_batches[w * per : (w + 1) * per] for w in range(num_train_workers)]
