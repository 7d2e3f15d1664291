"""Packed micro-batch layout (`MicroBatch`, src/prime_rl/trainer/rl/data.py:13-23) and deterministic
synthetic generators for it (the reference's `FakeDataLoader`, rl/data.py:26-54, plus the Qwen3-shaped
regimes of SURVEY.md §8d)."""

from __future__ import annotations

from typing import TypedDict

import torch


class MicroBatch(TypedDict):
    input_ids: torch.Tensor      # int64 [1, T]
    position_ids: torch.Tensor   # int64 [1, T], restarts at 0 per packed sample
    advantages: torch.Tensor     # fp32  [1, T]
    logprobs: torch.Tensor       # fp32  [1, T] inference log-probs (0 on prompt tokens)
    loss_mask: torch.Tensor      # bool  [1, T]
    temperature: float


def fake_micro_batch(seq_len: int, micro_batch_size: int = 1, vocab: int = 100, generator: torch.Generator | None = None) -> MicroBatch:
    """Same distribution as `FakeDataLoader._get_micro_batch` (rl/data.py:40-54), on the CPU."""
    T = micro_batch_size * seq_len
    return {
        "input_ids": torch.randint(0, vocab, (1, T), generator=generator),
        "position_ids": torch.cat([torch.arange(seq_len)] * micro_batch_size).unsqueeze(0),
        "advantages": torch.randn(T, generator=generator).unsqueeze(0),
        "logprobs": torch.randn(T, generator=generator).unsqueeze(0),
        "temperature": 1.0,
        "loss_mask": torch.ones(T, dtype=torch.bool).unsqueeze(0),
    }


def synthetic_micro_batch(
    segment_lengths: list[int],
    vocab: int,
    *,
    generator: torch.Generator,
    temperature: float = 1.0,
    prompt_frac: float = 0.0,
    multi_turn: bool = False,
    group_rewards: torch.Tensor | None = None,
) -> MicroBatch:
    """A packed row of `len(segment_lengths)` samples.  `prompt_frac` masks the first part of every
    sample (prompt tokens: mask 0, old log-prob 0, as `prepare_sample` does, orchestrator/batch.py:41-44);
    `multi_turn` additionally masks alternating "environment" runs inside the completion (wordle-style,
    orchestrator/config.py:463-468).  Advantages are one scalar per sample broadcast over its tokens."""
    T = sum(segment_lengths)
    ids = torch.randint(0, vocab, (T,), generator=generator)
    pos = torch.cat([torch.arange(n) for n in segment_lengths])
    mask = torch.ones(T, dtype=torch.bool)
    adv = torch.empty(T)
    o = 0
    for i, n in enumerate(segment_lengths):
        p = int(n * prompt_frac)
        mask[o : o + p] = False
        if multi_turn:
            turn = max(8, n // 16)
            k = o + p + turn
            while k < o + n:
                mask[k : min(k + turn // 2, o + n)] = False
                k += 2 * turn
        a = float(group_rewards[i]) if group_rewards is not None else float(torch.randn((), generator=generator))
        adv[o : o + n] = a
        o += n
    old = -torch.rand(T, generator=generator) * 2.0
    old[~mask] = 0.0
    return {
        "input_ids": ids.unsqueeze(0),
        "position_ids": pos.unsqueeze(0),
        "advantages": adv.unsqueeze(0),
        "logprobs": old.unsqueeze(0),
        "loss_mask": mask.unsqueeze(0),
        "temperature": float(temperature),
    }


def to_device(mb: MicroBatch, device) -> MicroBatch:
    """train.py:275-279."""
    out = dict(mb)
    for k in ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask"):
        out[k] = mb[k].to(device, non_blocking=True)
    return out  # type: ignore[return-value]
