"""GPU implementations behind the reference's loss API (src/prime_rl/trainer/rl/loss.py):
`compute_loss` (:61-124) as one segmented launch, and the materialised-logits helpers
`selective_log_softmax` (:13-39), `compute_entropy` (:43-47), `shift_logits` (:51-58)."""

from __future__ import annotations

from collections import OrderedDict
from typing import Any, Sequence

import torch

from . import _lib


# --------------------------------------------------------------------------- packed loss
def _pack(items: Sequence[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    """One contiguous 1-D tensor holding the items back to back.  The trainer passes
    `tensor.squeeze().split(lengths)` (rl/train.py:293-296), i.e. consecutive views of one storage:
    that case is recovered without a copy; independent tensors (the reference's unit tests) are
    concatenated."""
    first = items[0]
    same = first.dtype == dtype and first.dim() == 1
    if same:
        off = first.storage_offset()
        sp = first.untyped_storage().data_ptr()
        for t in items:
            if (t.dtype != dtype or t.dim() != 1 or t.untyped_storage().data_ptr() != sp or t.storage_offset() != off
                    or (t.numel() > 1 and t.stride(0) != 1)):
                same = False
                break
            off += t.numel()
        if same:
            total = off - first.storage_offset()
            return first.detach().as_strided((total,), (1,), first.storage_offset())
    return torch.cat([t.detach().reshape(-1).to(dtype) for t in items])


_CU_CACHE: "OrderedDict[tuple, tuple[torch.Tensor, torch.Tensor]]" = OrderedDict()
_CU_CACHE_ENTRIES = 256


def _cu_seqlens(lens: Sequence[int], dev: torch.device) -> torch.Tensor:
    """int32 [n + 1] segment bounds on the device.  The host copy is PINNED and the upload asynchronous (a pageable source would
    make the copy synchronous and cannot be captured in a CUDA graph); both tensors stay in a small LRU keyed by the lengths, so
    a repeated packing costs no copy at all and a copy captured in a graph keeps a live source."""
    key = (dev.type, dev.index, tuple(lens))
    hit = _CU_CACHE.get(key)
    if hit is not None:
        _CU_CACHE.move_to_end(key)
        return hit[1]
    bounds = [0]
    for ln in lens:
        bounds.append(bounds[-1] + ln)
    host = torch.tensor(bounds, dtype=torch.int32).pin_memory()
    cu = host.to(dev, non_blocking=True)
    _CU_CACHE[key] = (host, cu)
    if len(_CU_CACHE) > _CU_CACHE_ENTRIES:
        _CU_CACHE.popitem(last=False)
    return cu


def _loss_packed_raw(logp, old, adv, mask_u8, cu_seqlens, nseg, mode, length_norm, clip, loss_scale):
    lib = _lib.load()
    T = logp.numel()
    dev = logp.device
    n_out = T if mode == 0 else nseg
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ratio = torch.empty(n_out, dtype=torch.float32, device=dev)
    clipped = torch.empty(n_out, dtype=torch.float32, device=dev)
    is_clipped = torch.empty(n_out, dtype=torch.float32, device=dev)
    dlogp = torch.empty(T, dtype=torch.float32, device=dev)
    ws_bytes = lib.prl_loss_workspace_bytes(T, nseg)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rc = lib.prl_loss_packed_fwd(
        _lib.ptr(logp), _lib.ptr(old), _lib.ptr(adv), _lib.ptr(mask_u8), _lib.ptr(cu_seqlens), nseg, T, mode,
        1 if length_norm else 0, float(clip), float(loss_scale), _lib.ptr(loss), _lib.ptr(ratio), _lib.ptr(clipped),
        _lib.ptr(is_clipped), _lib.ptr(dlogp), _lib.ptr(ws), ws_bytes, _lib.stream(),
    )
    _lib.check(rc, "prl_loss_packed_fwd")
    return loss, ratio, clipped, is_clipped, dlogp


class _PackedLoss(torch.autograd.Function):
    """loss = f(logprobs); the forward kernel already produced d loss / d logprobs."""

    @staticmethod
    def forward(ctx, lens, old, adv, mask_u8, cu, mode, length_norm, clip, loss_scale, *logprob_items):
        logp = _pack(logprob_items, torch.float32)
        loss, ratio, clipped, is_clipped, dlogp = _loss_packed_raw(
            logp, old, adv, mask_u8, cu, len(lens), mode, length_norm, clip, loss_scale
        )
        ctx.save_for_backward(dlogp)
        ctx.lens = lens
        ctx.mark_non_differentiable(ratio, clipped, is_clipped)
        return loss, ratio, clipped, is_clipped

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        (dlogp,) = ctx.saved_tensors
        lib = _lib.load()
        out = torch.empty_like(dlogp)
        g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        rc = lib.prl_scale_f32(_lib.ptr(dlogp), _lib.ptr(g), _lib.ptr(out), dlogp.numel(), _lib.stream())
        _lib.check(rc, "prl_scale_f32")
        grads = out.split(ctx.lens)
        return (None,) * 9 + tuple(grads)


def compute_loss(
    logprobs: Any,
    old_logprobs: Any,
    advantages: Any,
    loss_mask: Any,
    loss_config: Any,
    loss_scale: int,
) -> tuple[torch.Tensor, dict[str, Any]]:
    """Drop-in for `compute_loss` (rl/loss.py:61-124): same arguments (sequences of 1-D tensors, one
    per packed sequence), same return value `(0-d loss, {"importance_ratio", "clipped_importance_ratio",
    "is_clipped"})`; the logging tensors have one entry per token in token mode and one per sequence in
    sequence mode.  The per-sequence Python loop becomes one segmented CUDA launch."""
    logprobs, old_logprobs = list(logprobs), list(old_logprobs)
    advantages, loss_mask = list(advantages), list(loss_mask)
    n = len(logprobs)
    if not (len(old_logprobs) == len(advantages) == len(loss_mask) == n):
        raise ValueError("compute_loss: the four sequences must have the same number of items")
    if loss_config.ratio_type not in _lib.PRL_RATIO:
        raise ValueError(f"unknown ratio_type {loss_config.ratio_type!r}")
    if n == 0:
        raise ValueError("compute_loss: empty input")
    _lib.require_cuda(*logprobs, *old_logprobs, *advantages, *loss_mask)
    lens = [int(t.numel()) for t in logprobs]
    for i in range(n):
        if not (old_logprobs[i].numel() == advantages[i].numel() == loss_mask[i].numel() == lens[i]):
            raise ValueError(f"compute_loss: item {i} has mismatched lengths")
    old = _pack(old_logprobs, torch.float32)
    adv = _pack(advantages, torch.float32)
    mask = _pack(loss_mask, torch.bool).view(torch.uint8)
    dev = old.device
    cu = _cu_seqlens(lens, dev)
    mode = _lib.PRL_RATIO[loss_config.ratio_type]
    loss, ratio, clipped, is_clipped = _PackedLoss.apply(
        tuple(lens), old, adv, mask, cu, mode, bool(loss_config.ratio_length_norm), float(loss_config.clip_ratio),
        float(loss_scale), *logprobs,
    )
    return loss, {"importance_ratio": ratio, "clipped_importance_ratio": clipped, "is_clipped": is_clipped}


# --------------------------------------------------------------------------- materialised-logits helpers
def _stats(logits: torch.Tensor, index: torch.Tensor | None, inv_temp: float, shift: bool, want_logp: bool, want_ent: bool):
    _lib.require_cuda(logits, index)
    lib = _lib.load()
    if logits.dim() != 3:
        raise ValueError("logits must be [batch, seq, vocab]")
    B, S, V = logits.shape
    x = logits.detach().contiguous().float()
    idx = None
    if index is not None:
        if tuple(index.shape) != (B, S):
            raise ValueError("index must be [batch, seq]")
        idx = index.contiguous().to(torch.int64)
    logp = torch.empty(B, S, dtype=torch.float32, device=x.device) if want_logp else None
    ent = torch.empty(B, S, dtype=torch.float32, device=x.device) if want_ent else None
    rc = lib.prl_logits_stats_f32(_lib.ptr(x), _lib.ptr(idx), B * S, S, V, float(inv_temp), 1 if shift else 0, _lib.ptr(logp),
                                  _lib.ptr(ent), _lib.stream())
    _lib.check(rc, "prl_logits_stats_f32")
    return logp, ent


class _LogitsLogProb(torch.autograd.Function):
    """log-prob of `index` under softmax(logits / tau) (optionally shifted) on MATERIALISED logits, differentiable w.r.t. the
    logits: forward = one read of [B,S,V], backward = one read + one write (the reference's autograd chain makes ~6 passes)."""

    @staticmethod
    def forward(ctx, logits, index, inv_temp, shift):
        logp, _ = _stats(logits, index, inv_temp, shift, True, False)
        ctx.save_for_backward(logits, index, logp)
        ctx.cfg = (inv_temp, shift)
        return logp

    @staticmethod
    def backward(ctx, g):
        logits, index, logp = ctx.saved_tensors
        inv_temp, shift = ctx.cfg
        lib = _lib.load()
        B, S, V = logits.shape
        x = logits.detach().contiguous().float()
        idx = index.contiguous().to(torch.int64)
        gg = g.detach().contiguous().float()
        d = torch.empty(B, S, V, dtype=torch.float32, device=x.device)
        rc = lib.prl_logits_logsoftmax_bwd_f32(_lib.ptr(x), _lib.ptr(idx), _lib.ptr(logp), _lib.ptr(gg), B * S, S, V, float(inv_temp),
                                               1 if shift else 0, _lib.ptr(d), _lib.stream())
        _lib.check(rc, "prl_logits_logsoftmax_bwd_f32")
        return d.to(logits.dtype), None, None, None


def selective_log_softmax(logits: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """`log_softmax(logits)[index]` in one pass over materialised logits (rl/loss.py:13-39; fp32, bf16 and fp16 inputs, the
    statistics are always fp32).  Differentiable w.r.t. `logits` (one more pass in the backward), so the reference trainer's
    `selective_log_softmax(shift_logits(logits) / temperature, input_ids)` keeps working unchanged; the fast path that
    never materialises fp32 logits is `fused_lm_head_logprobs_entropy` (lm_head.py)."""
    if logits.requires_grad and torch.is_grad_enabled():
        _lib.require_cuda(logits, index)
        if logits.dim() != 3 or tuple(index.shape) != tuple(logits.shape[:2]):
            raise ValueError("logits must be [batch, seq, vocab] and index [batch, seq]")
        return _LogitsLogProb.apply(logits, index, 1.0, False)
    return _stats(logits, index, 1.0, False, True, False)[0]


def compute_entropy(shifted_logits: torch.Tensor) -> torch.Tensor:
    """Entropy of softmax(logits) per position (rl/loss.py:43-47)."""
    return _stats(shifted_logits, None, 1.0, False, False, True)[1]


def shifted_logprobs_entropy(logits: torch.Tensor, input_ids: torch.Tensor, temperature: float):
    """rl/train.py:286-288 + :302-303 on materialised logits in ONE pass: shift, / temperature,
    log-prob of `input_ids`, entropy.  Returns (logprobs[B,S], entropy[B,S])."""
    if logits.requires_grad and torch.is_grad_enabled():  # log-probs differentiable, entropy detached (train.py:302-303)
        _lib.require_cuda(logits, input_ids)
        if logits.dim() != 3 or tuple(input_ids.shape) != tuple(logits.shape[:2]):
            raise ValueError("logits must be [batch, seq, vocab] and input_ids [batch, seq]")
        logp = _LogitsLogProb.apply(logits, input_ids, 1.0 / float(temperature), True)
        return logp, _stats(logits, None, 1.0 / float(temperature), True, False, True)[1]
    return _stats(logits, input_ids, 1.0 / float(temperature), True, True, True)


def shift_logits(logits: torch.Tensor) -> torch.Tensor:
    """Removes the final position and prepends a zero row (rl/loss.py:51-58).  Pure data movement,
    kept for API compatibility; the fused path never builds this tensor."""
    B, S, V = logits.shape
    out = torch.empty_like(logits)
    out[:, 0].zero_()
    out[:, 1:].copy_(logits[:, :-1])
    return out
