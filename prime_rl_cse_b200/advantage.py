"""Group-relative advantages on the GPU behind the reference's `compute_advantages` signature
(src/prime_rl/orchestrator/advantage.py:71-103)."""

from __future__ import annotations

from typing import Literal

import torch

from . import _lib

AdvantageType = Literal["drgrpo", "drgrpo-negclipped", "rloo", "opo", "grpo"]
REGISTRY = tuple(_lib.PRL_ADV_KINDS)  # the reference's keys (advantage.py:63-68) + "grpo" = (r - mean) / (std + 1e-6), an extension


def compute_advantages_tensor(
    rewards: torch.Tensor,
    completion_lengths: torch.Tensor | None,
    samples_per_problem: int,
    advantage_type: AdvantageType,
    token_offsets: torch.Tensor | None = None,
) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor]:
    """Device-resident form: rewards fp32 [N] (cuda), completion_lengths int64 [N] (cuda, needed for "opo").
    With `token_offsets` (int64 [N+1], cuda) also returns the per-token broadcast
    (`prepare_sample`, src/prime_rl/orchestrator/batch.py:44) as a second tensor of token_offsets[-1] floats."""
    if advantage_type not in _lib.PRL_ADV_KINDS:
        raise KeyError(advantage_type)  # REGISTRY[advantage_type], advantage.py:95
    _lib.require_cuda(rewards, completion_lengths, token_offsets)
    lib = _lib.load()
    rewards = rewards.contiguous().float()
    n = rewards.numel()
    assert n % samples_per_problem == 0  # advantage.py:90
    if completion_lengths is not None:
        completion_lengths = completion_lengths.contiguous().to(torch.int64)
    out = torch.empty_like(rewards)
    token_out = None
    if token_offsets is not None:
        token_offsets = token_offsets.contiguous().to(torch.int64)
        total = int(token_offsets[-1].item())
        token_out = torch.empty(total, dtype=torch.float32, device=rewards.device)
    rc = lib.prl_advantages_f32(
        _lib.ptr(rewards), _lib.ptr(completion_lengths), n, samples_per_problem, _lib.PRL_ADV_KINDS[advantage_type],
        _lib.ptr(out), _lib.ptr(token_offsets), _lib.ptr(token_out), _lib.stream(),
    )
    _lib.check(rc, "prl_advantages_f32")
    return out if token_out is None else (out, token_out)


def compute_advantages(
    rewards: list[float],
    completion_lengths: list[int],
    samples_per_problem: int,
    advantage_type: AdvantageType,
) -> list[float]:
    """Drop-in for the reference function: python lists in, python list out (fp32 values as doubles)."""
    assert len(rewards) % samples_per_problem == 0  # advantage.py:90
    if advantage_type not in _lib.PRL_ADV_KINDS:
        raise KeyError(advantage_type)
    if len(rewards) == 0:
        return []
    dev = torch.device("cuda", torch.cuda.current_device())
    r = torch.tensor(rewards, dtype=torch.float32).pin_memory().to(dev, non_blocking=True)
    ln = torch.tensor(completion_lengths, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
    adv = compute_advantages_tensor(r, ln, samples_per_problem, advantage_type)
    out = adv.cpu().tolist()
    assert len(rewards) == len(out)  # advantage.py:102
    return out
