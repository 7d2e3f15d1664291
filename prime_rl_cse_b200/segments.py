"""Segment boundaries of a packed row on the GPU (replaces `get_response_lengths`,
src/prime_rl/trainer/utils.py:36-69, whose host loop costs O(T) device syncs)."""

from __future__ import annotations

import torch

from . import _lib


def segments_from_position_ids(position_ids: torch.Tensor, max_segments: int | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns (cu_seqlens int32 [max_segments+1], nseg int32 [1]) on the device; no host sync."""
    _lib.require_cuda(position_ids)
    lib = _lib.load()
    pos = position_ids.reshape(-1).contiguous().to(torch.int64)
    T = pos.numel()
    if max_segments is None:
        max_segments = max(T // 2 + 1, 1)
    buf = torch.empty(max_segments + 2, dtype=torch.int32, device=pos.device)  # one buffer: the count sits right behind the bounds
    cu, nseg = buf[: max_segments + 1], buf[max_segments + 1 :]
    rc = lib.prl_segments_from_position_ids(_lib.ptr(pos), T, _lib.ptr(cu), _lib.ptr(nseg), max_segments, _lib.stream())
    _lib.check(rc, "prl_segments_from_position_ids")
    return cu, nseg


def get_response_lengths(position_ids: torch.Tensor) -> list[int]:
    """Drop-in for the reference function: python list of segment lengths (one D2H copy, one sync)."""
    if position_ids.numel() == 0:
        return [0]
    T = position_ids.numel()
    # a boundary needs pos[i-1] != 0 and pos[i+1] == 1, so there are at most T // 2 + 1 segments; rows with more than 4095 of them
    # (never seen: a packed row holds a handful of rollouts) take a second, full-size pass
    for cap in (min(T // 2 + 1, 4095), T // 2 + 1):
        cu, nseg = segments_from_position_ids(position_ids, cap)
        table = torch.cat([nseg, cu])  # views of one device buffer -> ONE device-to-host copy, one synchronisation
        host = torch.empty(table.shape, dtype=torch.int32, pin_memory=True)
        host.copy_(table, non_blocking=True)
        torch.cuda.current_stream(table.device).synchronize()
        b = host.tolist()
        n = b[0]
        if n >= 0:
            return [b[i + 2] - b[i + 1] for i in range(n)]
        if cap == T // 2 + 1:
            break
    raise RuntimeError("segment table overflow")


def response_lengths_host(position_ids: torch.Tensor) -> list[int]:
    """The same segment lengths from a HOST copy of position_ids (a data loader has one before it uploads the micro-batch):
    boundary at i iff pos[i] == 0 and pos[i-1] != 0 and i + 1 < T and pos[i+1] == 1 (trainer/utils.py:54-60), vectorised.
    Lets `rl_micro_step` run without any device-to-host synchronisation (`micro_batch["response_lengths"]`)."""
    if position_ids.is_cuda:
        raise ValueError("response_lengths_host works on a CPU tensor; use get_response_lengths for device tensors")
    p = position_ids.reshape(-1)
    T = p.numel()
    if T == 0:
        return [0]
    if T < 3:
        return [T]
    is_b = (p[1:-1] == 0) & (p[:-2] != 0) & (p[2:] == 1)
    b = [0] + (is_b.nonzero().reshape(-1) + 1).tolist() + [T]
    return [b[k + 1] - b[k] for k in range(len(b) - 1)]
