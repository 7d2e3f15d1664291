"""prime_rl_cse_b200 — B200-native (sm_100a) implementation of PRIME-RL's per-token RL loss hot path.

Drop-in surface (same names and call signatures as the reference):
  compute_advantages                      <- prime_rl.orchestrator.advantage
  compute_loss, selective_log_softmax,
  compute_entropy, shift_logits           <- prime_rl.trainer.rl.loss
  get_response_lengths                    <- prime_rl.trainer.utils
  LossConfig                              <- prime_rl.trainer.rl.config
plus the fused fast path `fused_lm_head_logprobs_entropy`, `FusedLMHead`, `rl_micro_step`.

Everything runs on CUDA through the C-ABI library `libprl_b200.so`; there is no CPU fallback.
"""

from .advantage import compute_advantages, compute_advantages_tensor
from .batch import first_fit_decreasing, prepare_batch
from .config import LossConfig
from .data import MicroBatch, fake_micro_batch, synthetic_micro_batch, to_device
from .lm_head import (FusedLMHead, fused_lm_head_logprobs_entropy, lm_head_backward_dw_scatter_raw, lm_head_backward_raw,
                      lm_head_backward_saved_raw,
                      lm_head_forward_raw, lm_head_forward_save_raw, saved_logits_policy)
from .loss import compute_entropy, compute_loss, selective_log_softmax, shift_logits, shifted_logprobs_entropy
from .extras import fused_linear_cross_entropy, recompute_logprobs
from .micro_step import rl_micro_step
from .rollout_io import DataLoader, FakeDataLoader, save_step, upload_micro_batches
from .segments import get_response_lengths, response_lengths_host, segments_from_position_ids
from .stats import TensorStats
from .parallel import (VocabShardedLMHead, exchange_partials, local_loss_scale, shard_micro_batches,
                       vocab_parallel_lm_head_logprobs_entropy, vocab_shard_bounds)
from .peer import PeerBuffer, ShardedGradAccumulator
from .rows import partition_positions

__all__ = [
    "compute_advantages", "compute_advantages_tensor", "LossConfig", "MicroBatch", "fake_micro_batch",
    "synthetic_micro_batch", "to_device", "FusedLMHead", "fused_lm_head_logprobs_entropy", "lm_head_backward_raw",
    "lm_head_forward_raw", "lm_head_forward_save_raw", "lm_head_backward_saved_raw", "saved_logits_policy", "compute_entropy", "compute_loss", "selective_log_softmax", "shift_logits",
    "shifted_logprobs_entropy", "rl_micro_step", "get_response_lengths", "segments_from_position_ids",
    "fused_linear_cross_entropy", "recompute_logprobs", "TensorStats", "prepare_batch", "first_fit_decreasing",
    "DataLoader", "FakeDataLoader", "save_step", "upload_micro_batches", "exchange_partials", "local_loss_scale", "shard_micro_batches", "vocab_parallel_lm_head_logprobs_entropy", "vocab_shard_bounds",
    "VocabShardedLMHead", "partition_positions", "response_lengths_host", "PeerBuffer", "ShardedGradAccumulator",
    "lm_head_backward_dw_scatter_raw",
]
