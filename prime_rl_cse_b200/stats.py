"""On-device logging statistics ("next" row N1 of SURVEY.md §8f).

The reference appends, per micro-batch, `tensor[loss_mask].to("cpu")` for every logged key (probs, old_probs, entropy,
importance ratios ...; src/prime_rl/trainer/rl/train.py:313-333) and at the end of the step all-gathers every masked
token across ranks to take mean/median/std/min/max (`Tensors.compute_stats`, src/prime_rl/trainer/utils.py:228-256).
`TensorStats` keeps five fp64 moments per key on the device, updated by one kernel per (key, micro-batch), and
all-reduces 5 scalars per key: no boolean indexing, no blocking D2H copy per micro-batch, no padded all-gather.
The median needs the values themselves and is therefore not reproduced (documented difference).
"""

from __future__ import annotations

import math

import torch

from . import _lib


class TensorStats:
    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.acc: dict[str, torch.Tensor] = {}
        lib = _lib.load()
        self._ws = torch.empty(lib.prl_masked_moments_workspace_bytes(), dtype=torch.uint8, device=self.device)

    def add(self, key: str, values: torch.Tensor, mask: torch.Tensor | None = None, *, exp: bool = False) -> None:
        """Accumulate f(values)[mask] into `key` (f = exp when `exp`, e.g. probs = exp(logprobs))."""
        _lib.require_cuda(values, mask)
        lib = _lib.load()
        x = values.detach().reshape(-1).contiguous().float()
        m = None
        if mask is not None:
            m = mask.reshape(-1).contiguous()
            if m.numel() != x.numel():
                raise ValueError("mask and values disagree")
            m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
        if key not in self.acc:
            self.acc[key] = torch.tensor([0.0, 0.0, 0.0, math.inf, -math.inf], dtype=torch.float64, device=self.device)
        rc = lib.prl_masked_moments_f32(_lib.ptr(x), _lib.ptr(m), x.numel(), 1 if exp else 0, _lib.ptr(self.acc[key]), _lib.ptr(self._ws),
                                        self._ws.numel(), _lib.stream())
        _lib.check(rc, "prl_masked_moments_f32")

    def compute_stats(self, group=None) -> dict[str, float]:
        """`{key}/mean|std|min|max|count` over all accumulated elements of all ranks; clears the accumulators.
        std is the unbiased estimator, as `Tensor.std()` in the reference (trainer/utils.py:249)."""
        import torch.distributed as dist

        out: dict[str, float] = {}
        keys = sorted(self.acc)
        if not keys:
            return out
        sums = torch.stack([self.acc[k][:3] for k in keys])
        mins = torch.stack([self.acc[k][3] for k in keys])
        maxs = torch.stack([self.acc[k][4] for k in keys])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(sums, group=group)
            dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
        host = torch.cat([sums, mins[:, None], maxs[:, None]], dim=1).cpu().tolist()  # ONE D2H copy for all keys
        for k, (n, s, q, mn, mx) in zip(keys, host):
            mean = s / n if n > 0 else float("nan")
            var = (q - s * s / n) / (n - 1) if n > 1 else float("nan")
            out[f"{k}/mean"], out[f"{k}/std"] = mean, math.sqrt(max(var, 0.0)) if n > 1 else float("nan")
            out[f"{k}/min"], out[f"{k}/max"], out[f"{k}/count"] = mn, mx, n
        self.acc.clear()
        return out
