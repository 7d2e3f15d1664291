"""On-device logging statistics ("next" row N1 of SURVEY.md §8f).

The reference appends, per micro-batch, `tensor[loss_mask].to("cpu")` for every logged key (probs, old_probs, entropy,
importance ratios ...; src/prime_rl/trainer/rl/train.py:313-333) and at the end of the step all-gathers every masked
token across ranks to take mean/median/std/min/max (`Tensors.compute_stats`, src/prime_rl/trainer/utils.py:228-256).
`TensorStats` keeps five fp64 moments per key on the device, updated by one kernel per (key, micro-batch), and
all-reduces 5 scalars per key: no boolean indexing, no blocking D2H copy per micro-batch, no padded all-gather.
The median (`torch.median`, the lower median) is exact too: every `add` appends order-preserving 32-bit keys of the masked
values to a device buffer (4 B per masked token), and `compute_stats` runs a 3-pass radix select in which the ranks
exchange three 2048-bin histograms per step for all keys together — still no gather of the values and no host sync.
"""

from __future__ import annotations

import math
import struct

import torch

from . import _lib


class TensorStats:
    _PASSES = ((21, 11), (10, 11), (0, 10))  # (shift, bits) of the radix-select passes, most significant first

    def __init__(self, device=None, median: bool = True):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.acc: dict[str, torch.Tensor] = {}
        self.median = median
        # per key: [keys uint32 buffer (as int32 storage), device counter (uint64 as int64 storage), elements offered so far]
        self.keys: dict[str, list] = {}
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
        if self.median and x.numel() > 0:
            if key not in self.keys:
                self.keys[key] = [torch.empty(max(4 * x.numel(), 1 << 16), dtype=torch.int32, device=self.device),
                                  torch.zeros(1, dtype=torch.int64, device=self.device), 0]
            buf, counter, offered = self.keys[key]
            if offered + x.numel() > buf.numel():  # grow geometrically (device copy, no sync; the true count stays on the device)
                bigger = torch.empty(max(2 * buf.numel(), offered + x.numel()), dtype=torch.int32, device=self.device)
                bigger[: buf.numel()].copy_(buf)
                self.keys[key][0] = buf = bigger
            rc = lib.prl_masked_keys_u32(_lib.ptr(x), _lib.ptr(m), x.numel(), 1 if exp else 0, _lib.ptr(buf), buf.numel(), _lib.ptr(counter),
                                         _lib.stream())
            _lib.check(rc, "prl_masked_keys_u32")
            self.keys[key][2] = offered + x.numel()

    def add_micro_step(self, out: dict, micro_batch: dict, ratio_type: str = "token", recomputed_logprob_error: torch.Tensor | None = None) -> None:
        """Everything rl/train.py:313-333 appends after a micro-step, under the same keys, without leaving the device:
        `out` is what `rl_micro_step` returns (loss, logprobs, entropy, importance_ratio, clipped_importance_ratio, is_clipped),
        `micro_batch` the MicroBatch it ran on.  probs = exp(logprobs)[mask], old_probs = exp(old logprobs)[mask], entropy[mask],
        recomputed_logprob_error[mask] (if given), loss, and the three loss tensors ([mask] in token mode, per sequence otherwise)."""
        mask = micro_batch["loss_mask"]
        self.add("probs", out["logprobs"], mask, exp=True)                       # train.py:313
        self.add("old_probs", micro_batch["logprobs"], mask, exp=True)           # train.py:314
        self.add("entropy", out["entropy"], mask)                                # train.py:315
        if recomputed_logprob_error is not None:
            self.add("recomputed_logprob_error", recomputed_logprob_error, mask)  # train.py:316-318
        self.add("loss", out["loss"].reshape(1))                                 # train.py:319
        for key in ("importance_ratio", "clipped_importance_ratio", "is_clipped"):
            if key in out:                                                       # train.py:328-333
                self.add(key, out[key], None if ratio_type == "sequence" else mask)

    def _medians(self, keys: list[str], counts: torch.Tensor, group) -> torch.Tensor:
        """[nkeys] int64 device tensor: the order-preserving key of the global lower median of every logged key (-1 if empty).
        `counts` = the all-reduced element counts (fp64)."""
        import torch.distributed as dist

        lib = _lib.load()
        nk = len(keys)
        state = torch.zeros(nk, 4, dtype=torch.int64, device=self.device)
        state[:, 2] = torch.floor((counts - 1.0) / 2.0).to(torch.int64)  # k = (n - 1) // 2, -1 for n == 0
        hist = torch.zeros(nk, 2048, dtype=torch.int64, device=self.device)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        for shift, bits in self._PASSES:
            for i, k in enumerate(keys):
                if k in self.keys:
                    buf, counter, offered = self.keys[k]
                    rc = lib.prl_radix_select_hist(_lib.ptr(buf), _lib.ptr(counter), min(offered, buf.numel()), _lib.ptr(state[i]), shift, bits,
                                                   _lib.ptr(hist[i]), _lib.stream())
                    _lib.check(rc, "prl_radix_select_hist")
            if multi:
                dist.all_reduce(hist, group=group)
            for i in range(nk):
                _lib.check(lib.prl_radix_select_pick(_lib.ptr(hist[i]), _lib.ptr(state[i]), shift, bits, _lib.stream()), "prl_radix_select_pick")
        return torch.where(state[:, 2] >= 0, state[:, 0], torch.full_like(state[:, 0], -1))

    @staticmethod
    def _key_to_float(key: int) -> float:
        if key < 0:
            return float("nan")
        u = (key ^ 0x80000000) if (key & 0x80000000) else (~key & 0xFFFFFFFF)
        return struct.unpack("<f", struct.pack("<I", u))[0]

    def compute_stats(self, group=None) -> dict[str, float]:
        """`{key}/mean|median|std|min|max|count` over all accumulated elements of all ranks; clears the accumulators.
        std is the unbiased estimator, as `Tensor.std()` in the reference (trainer/utils.py:249); median is `torch.median`'s
        lower median (utils.py:248), nan if any value is nan (as in torch; detected through the sum of squares, which +-inf
        cannot turn into nan) or nothing was logged."""
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
        cols = [sums, mins[:, None], maxs[:, None]]
        if self.median:
            cols.append(self._medians(keys, sums[:, 0], group).double()[:, None])  # keys < 2^32: exact in fp64
        host = torch.cat(cols, dim=1).cpu().tolist()  # ONE D2H copy for all keys
        for k, row in zip(keys, host):
            n, s, q, mn, mx = row[:5]
            mean = s / n if n > 0 else float("nan")
            var = (q - s * s / n) / (n - 1) if n > 1 else float("nan")
            out[f"{k}/mean"], out[f"{k}/std"] = mean, math.sqrt(max(var, 0.0)) if n > 1 else float("nan")
            out[f"{k}/min"], out[f"{k}/max"], out[f"{k}/count"] = mn, mx, n
            if self.median:
                out[f"{k}/median"] = float("nan") if (n <= 0 or math.isnan(q)) else self._key_to_float(int(row[5]))
        self.acc.clear()
        self.keys.clear()
        return out
