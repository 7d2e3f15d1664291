""""Next" rows of SURVEY.md §8(f) that reuse the fused LM-head kernels unchanged.

N2  forward-only recompute of the old log-probs (src/prime_rl/trainer/rl/train.py:216-253)
N3  SFT cross-entropy on the same contraction (src/prime_rl/trainer/sft/train.py:209-216)
"""

from __future__ import annotations

import torch

from . import _lib
from .lm_head import _FusedLMHead, lm_head_forward_raw


@torch.no_grad()
def recompute_logprobs(hidden: torch.Tensor, weight: torch.Tensor, input_ids: torch.Tensor, temperature: float,
                       old_logprobs: torch.Tensor | None = None):
    """train.py:236-242 without materialising logits: `recomputed = selective_log_softmax(shift_logits(logits) /
    temperature, input_ids)` and, if `old_logprobs` is given, `recomputed_logprob_error = exp(recomputed - old)`.
    hidden [..., T, H] bf16, weight [V, H] bf16, input_ids [..., T].  Returns (logprobs[..., T], error | None)."""
    _lib.require_cuda(hidden, weight, input_ids, old_logprobs)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("recompute_logprobs needs bf16 hidden and weight (an fp32 logprob model must be cast explicitly)")
    lead = input_ids.shape
    h2 = hidden.reshape(-1, hidden.shape[-1]).contiguous()
    i2 = input_ids.reshape(-1).contiguous().to(torch.int64)
    logp, _, _, _ = lm_head_forward_raw(h2, weight.contiguous(), i2, 1.0 / float(temperature), shift=True)
    logp = logp.reshape(lead)
    err = torch.exp(logp - old_logprobs) if old_logprobs is not None else None
    return logp, err


def fused_linear_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, target_ids: torch.Tensor, loss_mask: torch.Tensor,
                               *, round_bf16: bool = True) -> torch.Tensor:
    """SFT loss of the reference (sft/train.py:209-216): `cross_entropy(lm_head(hidden), target_ids,
    reduction="none")[loss_mask].mean()` — targets are already shifted there, there is no temperature and no
    `shift_logits`.  hidden [..., T, H] bf16, weight [V, H] bf16, target_ids / loss_mask [..., T].  Differentiable w.r.t.
    hidden and weight; the fp32 [T, V] logits are never materialised (one bf16 copy is kept for the backward when it fits)."""
    _lib.require_cuda(hidden, weight, target_ids, loss_mask)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("fused_linear_cross_entropy needs bf16 hidden and weight")
    if hidden.shape[:-1] != target_ids.shape or target_ids.shape != loss_mask.shape:
        raise ValueError("hidden [..., T, H], target_ids [..., T] and loss_mask [..., T] disagree")
    h2 = hidden.reshape(1, -1, hidden.shape[-1])
    logp, _ = _FusedLMHead.apply(h2, weight, target_ids.reshape(1, -1), 1.0, False, bool(round_bf16), None, None, False)
    mask = loss_mask.reshape(1, -1)
    n = mask.sum()
    return -(logp * mask).sum() / n  # mean over unmasked tokens (nan for an all-False mask, like the reference)
