"""`LossConfig` — same fields, defaults and validation as the reference's pydantic model
(src/prime_rl/trainer/rl/config.py:19-27).  `compute_loss` only duck-types `.ratio_type`,
`.ratio_length_norm`, `.clip_ratio`, so the reference's own object can be passed instead."""

from __future__ import annotations

from dataclasses import dataclass


@dataclass
class LossConfig:
    ratio_type: str = "token"  # "token" | "sequence"
    ratio_length_norm: bool = False
    clip_ratio: float = 8.0

    def __post_init__(self):
        if self.ratio_type not in ("token", "sequence"):
            raise ValueError("ratio_type must be 'token' or 'sequence'")
        if self.clip_ratio < 0:
            raise ValueError("clip_ratio must be >= 0")
