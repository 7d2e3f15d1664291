"""Masked positions of a packed row on the device (include/prl_b200.h "Masked positions").

`loss_mask == False` tokens (prompts, environment turns: src/prime_rl/orchestrator/batch.py:41-44 of the reference) have
d loss / d logp = 0 (src/prime_rl/trainer/rl/loss.py:104), so the LM-head backward only has to visit the ACTIVE positions.
The row is partitioned (active positions first) by one kernel, `hidden` is gathered into that order, per-token outputs and
d_hidden are scattered back -- all by this library's own kernels; no nonzero / index_select / index_copy_ on the path."""

from __future__ import annotations

import math

import torch

from . import _lib


class Partition:
    """order[j] = position in slot j (active first), slot[t] = slot of position t (-1 for t < shift), n_active (host int,
    fetched lazily: one 4-byte D2H copy)."""

    __slots__ = ("order", "slot", "_n_dev", "_n", "T", "shift")

    def __init__(self, order, slot, n_dev, T, shift, n_active=None):
        self.order, self.slot, self._n_dev, self._n, self.T, self.shift = order, slot, n_dev, n_active, T, shift

    @property
    def n_active(self) -> int:
        if self._n is None:
            self._n = int(self._n_dev.item())
        return self._n

    @property
    def rows(self) -> int:
        """source rows that feed an output position"""
        return max(self.T - self.shift, 0)


def partition_positions(loss_mask: torch.Tensor, shift: bool = True, n_active: int | None = None) -> Partition:
    """Launches the partition kernel on the current stream and returns without synchronising (`n_active` syncs when read,
    unless the caller already knows it -- e.g. a data loader that counted `loss_mask[shift:]` on the host copy)."""
    _lib.require_cuda(loss_mask)
    m = loss_mask.detach().reshape(-1)
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    elif m.dtype != torch.uint8:
        raise ValueError("loss_mask must be a bool (or uint8) tensor")
    m = m.contiguous()
    T = m.numel()
    dev = m.device
    order = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    slot = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    n_dev = torch.empty(1, dtype=torch.int32, device=dev)
    rc = _lib.load().prl_partition_positions(_lib.ptr(m), T, 1 if shift else 0, _lib.ptr(order), _lib.ptr(slot), _lib.ptr(n_dev),
                                             _lib.stream())
    _lib.check(rc, "prl_partition_positions")
    return Partition(order, slot, n_dev, T, 1 if shift else 0, None if n_active is None else int(n_active))


def gather_rows(hidden: torch.Tensor, ids: torch.Tensor | None, part: Partition, n: int):
    """(hidden[order[:n] - shift], ids[order[:n]]): the first n slots of the partitioned row."""
    H = hidden.shape[1]
    out = torch.empty(n, H, dtype=torch.bfloat16, device=hidden.device)
    ids_out = torch.empty(n, dtype=torch.int64, device=hidden.device) if ids is not None else None
    rc = _lib.load().prl_gather_rows_bf16(_lib.ptr(hidden), _lib.ptr(ids), _lib.ptr(part.order), n, part.shift, H, _lib.ptr(out),
                                          _lib.ptr(ids_out), _lib.stream())
    _lib.check(rc, "prl_gather_rows_bf16")
    return out, ids_out


def gather_tokens(part: Partition, n: int, *vectors: torch.Tensor) -> list[torch.Tensor]:
    """x[order[:n]] for up to three fp32 [T] vectors."""
    if not 1 <= len(vectors) <= 3:
        raise ValueError("one to three vectors")
    outs = [torch.empty(n, dtype=torch.float32, device=v.device) for v in vectors]
    args = []
    for v, o in zip(vectors, outs):
        args += [_lib.ptr(v), _lib.ptr(o)]
    args += [_lib.ptr(None)] * (6 - len(args))
    rc = _lib.load().prl_gather_tokens_f32(_lib.ptr(part.order), n, *args, _lib.stream())
    _lib.check(rc, "prl_gather_tokens_f32")
    return outs


def scatter_tokens(part: Partition, n_valid: int, *pairs) -> list[torch.Tensor]:
    """Back to position order.  pairs = (vector in slot order [>= n_valid], value for positions without a source row); positions
    whose slot is >= n_valid (not computed) get 0."""
    if not 1 <= len(pairs) <= 3:
        raise ValueError("one to three (vector, first_value) pairs")
    dev = part.slot.device
    outs = [torch.empty(part.T, dtype=torch.float32, device=dev) for _ in pairs]
    args = []
    for (v, first), o in zip(pairs, outs):
        args += [_lib.ptr(v), _lib.ptr(o), float(first)]
    for _ in range(3 - len(pairs)):
        args += [_lib.ptr(None), _lib.ptr(None), 0.0]
    rc = _lib.load().prl_scatter_tokens_f32(_lib.ptr(part.slot), part.T, n_valid, *args, _lib.stream())
    _lib.check(rc, "prl_scatter_tokens_f32")
    return outs


def scatter_rows(rows_c: torch.Tensor | None, part: Partition, n_valid: int, H: int) -> torch.Tensor:
    """d_hidden [T, H] bf16 in source-row order from the gradient of the first n_valid slots; zeros elsewhere."""
    out = torch.empty(part.T, H, dtype=torch.bfloat16, device=part.slot.device)
    rc = _lib.load().prl_scatter_rows_bf16(_lib.ptr(rows_c), _lib.ptr(part.slot), part.T, n_valid, part.shift, H, _lib.ptr(out),
                                           _lib.stream())
    _lib.check(rc, "prl_scatter_rows_bf16")
    return out


def first_row_values(V: int) -> tuple[float, float, float]:
    """(logp, entropy, lse) of shift_logits' all-zero row (rl/loss.py:55-56): uniform over V"""
    ln_v = math.log(float(V))
    return -ln_v, ln_v, ln_v
