"""One RL micro-step of the hot path, mirroring src/prime_rl/trainer/rl/train.py:275-310 of the
reference with the fused kernels: segment detection, fused LM head (log-probs + entropy), packed
clipped-ratio loss, backward to d_hidden / d_W_lm."""

from __future__ import annotations

from typing import Any

import torch

from . import rows
from .lm_head import fused_lm_head_logprobs_entropy
from .loss import compute_loss
from .segments import get_response_lengths


def rl_micro_step(
    hidden: torch.Tensor,
    lm_head_weight: torch.Tensor,
    micro_batch: dict[str, Any],
    loss_config: Any,
    loss_scale: int,
    *,
    backward: bool = True,
    main_grad: torch.Tensor | None = None,
    response_lengths: list[int] | None = None,
    skip_masked: bool = True,
    forward_skips_masked: bool = False,
    save_logits: bool | None = None,
) -> dict[str, torch.Tensor]:
    """`hidden` [1,T,H] bf16 is what the transformer body hands to `lm_head` (requires_grad as the
    caller wishes); `micro_batch` is a `MicroBatch` (rl/data.py:13-23) whose tensors are already on the
    device (train.py:275-279).  Returns loss, logprobs, entropy and the three logging tensors of
    `compute_loss`; with `backward=True` also runs `loss.backward()` (train.py:310).  `skip_masked`: tokens with
    `loss_mask == False` get a zero gradient (loss.py:104), so the backward only visits the unmasked positions (the row is
    partitioned "active first" on the device, rows.py); a no-op when every token is unmasked.  `forward_skips_masked` (opt-in) also restricts the FORWARD to the
    unmasked positions: log-probs / entropy come back as 0 at masked positions (the reference computes values there that
    nothing reads: the loss multiplies by the mask and the trainer logs `[loss_mask]` slices, train.py:313-333).
    `save_logits`: see `fused_lm_head_logprobs_entropy` (None / False = recompute backward, True = keep the bf16 logits).
    Optional micro-batch keys `response_lengths` (list[int]) and `n_active` (int: unmasked positions from position 1 on), as
    `rollout_io.upload_micro_batches` attaches them from the host copy, remove the step's two device-to-host syncs."""
    input_ids = micro_batch["input_ids"]
    position_ids = micro_batch["position_ids"]
    advantages = micro_batch["advantages"]
    loss_mask = micro_batch["loss_mask"]
    old_logprobs = micro_batch["logprobs"]
    temperature = micro_batch["temperature"]

    # the partition kernel is queued before the segment lengths are fetched, so the one D2H sync of train.py:291 also
    # covers its 4-byte active-position count.  A loader that saw the micro-batch on the host (rollout_io.upload_micro_batches)
    # attaches "response_lengths" and "n_active": then the whole micro-step runs without a device-to-host synchronisation.
    partition = None
    if (skip_masked and backward) or forward_skips_masked:
        partition = rows.partition_positions(loss_mask, shift=True, n_active=micro_batch.get("n_active"))

    # train.py:291 -- done first so its single D2H sync happens before the heavy kernels are queued
    if response_lengths is None:
        response_lengths = micro_batch.get("response_lengths")
    if response_lengths is None:
        response_lengths = get_response_lengths(position_ids)

    # train.py:285-288 + 302-303
    logprobs, entropy = fused_lm_head_logprobs_entropy(hidden, lm_head_weight, input_ids, temperature, main_grad=main_grad,
                                                       loss_mask=partition, forward_active_only=forward_skips_masked,
                                                       save_logits=save_logits)

    # train.py:292-299
    loss, loss_tensors = compute_loss(
        # `.reshape(-1)` == the reference's `.squeeze()` for a [1, T] row, and also works for T == 1
        logprobs=logprobs.reshape(-1).split(response_lengths),
        old_logprobs=old_logprobs.reshape(-1).split(response_lengths),
        advantages=advantages.reshape(-1).split(response_lengths),
        loss_mask=loss_mask.reshape(-1).split(response_lengths),
        loss_config=loss_config,
        loss_scale=loss_scale,
    )
    if backward:
        loss.backward()  # train.py:310
    out = {"loss": loss.detach(), "logprobs": logprobs.detach(), "entropy": entropy}
    out.update(loss_tensors)
    return out
