"""numpy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Nothing here is imported by the product package.

Conventions: `T` tokens of one packed row, `H` hidden, `V` vocab, `tau`
temperature.  All public functions take/return numpy arrays; `dtype` selects the
arithmetic type (np.float32 mirrors the reference, np.float64 is the tight
checker for tolerance-based GPU parity).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

__all__ = [
    "LossCfg",
    "ADVANTAGE_KINDS",
    "EXTENSION_ADVANTAGE_KINDS",
    "sum_l8",
    "advantages",
    "response_lengths",
    "bf16_round",
    "lm_head_logits",
    "shift_rows",
    "logprobs_entropy",
    "loss_fwd_bwd",
    "lm_head_backward",
    "micro_step",
    "shard_partials",
    "merge_partials",
    "sft_cross_entropy",
]

ADVANTAGE_KINDS = ("drgrpo", "drgrpo-negclipped", "rloo", "opo")
# NOT in the reference's registry (advantage.py:63-68 is Dr.GRPO: mean only).  BASELINE.json's north_star names "per-group
# reward mean/std", so the product also offers GRPO's std-normalised form; this is its definition:
#   a_i = (r_i - mean) / (std + 1e-6), std = sqrt(sum (r_i - mean)^2 / (G - 1))  (unbiased, as torch.std), L8-ordered sums.
EXTENSION_ADVANTAGE_KINDS = ("grpo",)
GRPO_EPS = np.float32(1e-6)


@dataclass
class LossCfg:
    """Mirror of `LossConfig` (src/prime_rl/trainer/rl/config.py:19-27)."""

    ratio_type: str = "token"
    ratio_length_norm: bool = False
    clip_ratio: float = 8.0


# --------------------------------------------------------------------------- #
# advantages  (src/prime_rl/orchestrator/advantage.py:10-103)
# --------------------------------------------------------------------------- #
def sum_l8(x: np.ndarray) -> np.float32:
    """fp32 sum in the order the CUDA kernel uses (and ATen's CPU `sum` uses for
    G in {1..4, 8, 16, 24, 32} on the AVX512 build probed at survey time):
    8 strided lane accumulators acc[j] = x[j] + x[j+8] + ..., the lanes added
    left to right, then the tail (len % 8 elements) added left to right.
    """
    f = np.float32
    x = np.asarray(x, dtype=np.float32)
    n = (len(x) // 8) * 8
    if n == 0:
        s = f(0.0)
        for v in x:
            s = f(s + v)
        return s
    acc = [f(v) for v in x[:8]]
    for i in range(8, n):
        acc[i % 8] = f(acc[i % 8] + x[i])
    s = acc[0]
    for v in acc[1:]:
        s = f(s + v)
    for v in x[n:]:
        s = f(s + v)
    return s


def advantages(rewards: Sequence[float], lengths: Sequence[int], group: int, kind: str) -> np.ndarray:
    """Group-relative advantages, fp32.  Follows advantage.py:71-103 (grouping,
    :90-94) and the four registry functions: drgrpo :18, drgrpo-negclipped :32,
    rloo :41-43 (left-to-right `adv * G / (G - 1)`; G == 1 gives nan like the
    reference), opo :55-57.
    """
    if kind not in ADVANTAGE_KINDS + EXTENSION_ADVANTAGE_KINDS:
        raise KeyError(kind)
    r = np.asarray(rewards, dtype=np.float32)
    ln = np.asarray(lengths, dtype=np.int64)
    if len(r) % group != 0:
        raise AssertionError("len(rewards) % samples_per_problem != 0")  # advantage.py:90
    f = np.float32
    out = np.empty_like(r)
    for g0 in range(0, len(r), group):
        rg = r[g0 : g0 + group]
        if kind == "opo":
            w = ln[g0 : g0 + group].astype(np.float32)
            with np.errstate(divide="ignore", invalid="ignore"):
                base = f(sum_l8(rg * w) / sum_l8(w))
            a = (rg - base).astype(np.float32)
        else:
            mean = f(sum_l8(rg) / f(group))
            a = (rg - mean).astype(np.float32)
            if kind == "drgrpo-negclipped":
                a = np.maximum(a, f(0.0))
            elif kind == "rloo":
                with np.errstate(divide="ignore", invalid="ignore"):
                    a = ((a * f(group)).astype(np.float32) / f(group - 1)).astype(np.float32)
            elif kind == "grpo":
                with np.errstate(divide="ignore", invalid="ignore"):
                    std = np.sqrt(f(sum_l8((a * a).astype(np.float32)) / f(group - 1)), dtype=np.float32)
                    a = (a / f(std + GRPO_EPS)).astype(np.float32)
        out[g0 : g0 + group] = a
    return out


# --------------------------------------------------------------------------- #
# segment boundaries  (src/prime_rl/trainer/utils.py:36-69)
# --------------------------------------------------------------------------- #
def response_lengths(position_ids: np.ndarray) -> list[int]:
    """Boundary at i (1 <= i) iff pos[i] == 0 and pos[i-1] != 0 and i+1 < T and
    pos[i+1] == 1 (utils.py:54-60); lengths are differences of boundaries with T
    appended (:63-67)."""
    p = np.asarray(position_ids).reshape(-1)
    T = len(p)
    bounds = [0]
    for i in range(1, T):
        if p[i] == 0 and p[i - 1] != 0 and i + 1 < T and p[i + 1] == 1:
            bounds.append(i)
    bounds.append(T)
    return [bounds[i + 1] - bounds[i] for i in range(len(bounds) - 1)]


# --------------------------------------------------------------------------- #
# LM head -> shifted, temperature-scaled log-probs and entropy
# --------------------------------------------------------------------------- #
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even), returned as fp32.  This is
    what `nn.Linear` in bf16 followed by `.float()` does to the logits
    (src/prime_rl/trainer/model.py:254 + rl/train.py:285)."""
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    u = x32.view(np.uint32).astype(np.uint64)
    rnd = ((u >> 16) & 1) + 0x7FFF
    u = ((u + rnd) & 0xFFFF0000).astype(np.uint32)
    return u.view(np.float32).reshape(x32.shape)


def lm_head_logits(hidden: np.ndarray, weight: np.ndarray, round_bf16: bool, dtype=np.float64) -> np.ndarray:
    """logits[T,V] = hidden[T,H] @ weight[V,H]^T (model.py:250-254, HF lm_head
    is a bias-free nn.Linear).  With `round_bf16` the accumulator is rounded to
    bf16 as the GPU reference's bf16 GEMM output is."""
    acc = hidden.astype(dtype) @ weight.astype(dtype).T
    if round_bf16:
        return bf16_round(acc.astype(np.float32)).astype(dtype)
    return acc


def shift_rows(logits: np.ndarray) -> np.ndarray:
    """`shift_logits` (rl/loss.py:51-58) on one packed row: drop the last
    position, prepend an all-zero row."""
    out = np.zeros_like(logits)
    out[1:] = logits[:-1]
    return out


def logprobs_entropy(shifted: np.ndarray, ids: np.ndarray, tau: float, dtype=np.float64):
    """z = shifted / tau (rl/train.py:287); logp = z[ids] - logsumexp(z)
    (rl/loss.py:26-30); entropy = logsumexp(z) - sum(softmax(z) * z)
    (rl/loss.py:43-47).  Returns (logp[T], entropy[T], lse[T])."""
    z = shifted.astype(dtype) / dtype(tau)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    s = e.sum(axis=-1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    logp = z[np.arange(z.shape[0]), np.asarray(ids).reshape(-1)] - lse
    ent = lse - ((e / s) * z).sum(axis=-1)
    return logp, ent, lse


# --------------------------------------------------------------------------- #
# compute_loss  (src/prime_rl/trainer/rl/loss.py:61-124) + its gradient
# --------------------------------------------------------------------------- #
def loss_fwd_bwd(
    logp: np.ndarray,
    old: np.ndarray,
    adv: np.ndarray,
    mask: np.ndarray,
    lengths: Sequence[int],
    cfg: LossCfg,
    loss_scale: float,
    dtype=np.float32,
):
    """Per segment (loss.py:89): delta = logp - old (:90).  token mode: r =
    exp(delta) (:98), clipped = min(r, c) (:99), l = -clipped * A (:100),
    is_clipped = r > c (:101), seg = sum l[mask] (:104).  sequence mode
    (:92-96,107-108): L = sum delta[mask]; optional / max(n,1); clamp(max=10);
    r = exp(L); seg = sum(-min(r,c) * A)[mask] / max(n,1).  Total / max(S,1)
    (:118).  Also returns d loss / d logp (what autograd gives, SURVEY §8a row
    a10: clamp passes gradient at equality).

    Returns (loss, importance_ratio, clipped_importance_ratio, is_clipped, dlogp).
    In sequence mode the three tensors have one entry per segment."""
    logp = np.asarray(logp, dtype=dtype)
    old = np.asarray(old, dtype=dtype)
    adv = np.asarray(adv, dtype=dtype)
    mask = np.asarray(mask).astype(bool)
    c = dtype(cfg.clip_ratio)
    S = dtype(max(loss_scale, 1))
    total = dtype(0.0)
    ratios, clippeds, flags = [], [], []
    dlogp = np.zeros_like(logp)
    o = 0
    for n_tok in lengths:
        sl = slice(o, o + n_tok)
        o += n_tok
        d = logp[sl] - old[sl]
        mk = mask[sl]
        a = adv[sl]
        if cfg.ratio_type == "sequence":
            n = dtype(max(int(mk.sum()), 1))
            L = d[mk].sum(dtype=dtype)
            if cfg.ratio_length_norm:
                L = L / n
            pre = L
            L = min(L, dtype(10.0))
            r = np.exp(np.array([L], dtype=dtype))
            cl = np.minimum(r, c)
            seg = (-(cl) * a)[mk].sum(dtype=dtype) / n
            # d seg / d logp_t = -mask_t * (sum mask*A / n) * r * 1[r<=c] * 1[pre<=10] * (1/n if norm)
            coef = -(a[mk].sum(dtype=dtype) / n) * r[0]
            coef = coef * (1.0 if r[0] <= c else 0.0) * (1.0 if pre <= dtype(10.0) else 0.0)
            if cfg.ratio_length_norm:
                coef = coef / n
            dlogp[sl] = np.where(mk, coef, 0.0) / S
        else:
            r = np.exp(d)
            cl = np.minimum(r, c)
            seg = (-(cl) * a)[mk].sum(dtype=dtype)
            dlogp[sl] = np.where(mk & (r <= c), -a * r, 0.0) / S
        total = total + seg
        ratios.append(r)
        clippeds.append(cl)
        flags.append((r > c).astype(dtype))
    loss = total / S
    return (
        dtype(loss),
        np.concatenate(ratios).astype(dtype),
        np.concatenate(clippeds).astype(dtype),
        np.concatenate(flags).astype(dtype),
        dlogp.astype(dtype),
    )


# --------------------------------------------------------------------------- #
# backward of the LM-head chain  (autograd through rl/train.py:285-288)
# --------------------------------------------------------------------------- #
def lm_head_backward(
    hidden: np.ndarray,
    weight: np.ndarray,
    ids: np.ndarray,
    tau: float,
    lse: np.ndarray,
    dlogp: np.ndarray,
    round_bf16: bool,
    dtype=np.float64,
    logits: np.ndarray | None = None,
):
    """dz[t,v] = dlogp[t] * (1[v == ids[t]] - softmax(z[t])[v]); dlogits[t-1] =
    dz[t] / tau for t >= 1 (row 0 of the shifted tensor is a constant, the last
    un-shifted row is dropped: rl/loss.py:54-57); with `round_bf16` dlogits is
    rounded to bf16 (backward of `.float()` on bf16 logits, rl/train.py:285).
    dHidden = dlogits @ W, dW = dlogits^T @ hidden.  Returns (dHidden, dW).
    `logits` (optional, [T,V]): use these un-shifted logits instead of recomputing them (the "same-logits" checker:
    the GPU kernel's own bf16 logits block, so that accumulation-order effects of the GEMM are taken out)."""
    T = hidden.shape[0]
    logits = lm_head_logits(hidden, weight, round_bf16, dtype) if logits is None else np.asarray(logits, dtype=dtype)
    z = shift_rows(logits) / dtype(tau)
    p = np.exp(z - np.asarray(lse, dtype=dtype)[:, None])
    dz = -p * np.asarray(dlogp, dtype=dtype)[:, None]
    dz[np.arange(T), np.asarray(ids).reshape(-1)] += np.asarray(dlogp, dtype=dtype)
    dlogits = np.zeros_like(dz)
    dlogits[:-1] = dz[1:] / dtype(tau)
    if round_bf16:
        dlogits = bf16_round(dlogits.astype(np.float32)).astype(dtype)
    d_hidden = dlogits @ weight.astype(dtype)
    d_weight = dlogits.T @ hidden.astype(dtype)
    return d_hidden, d_weight


def micro_step(
    hidden, weight, ids, position_ids, old, adv, mask, tau, cfg: LossCfg, loss_scale, round_bf16=True, dtype=np.float64,
    logits=None,
):
    """The whole of rl/train.py:285-310 for one packed micro-batch, given the
    final hidden states: returns a dict with logp, entropy, lse, loss, the three
    logging tensors, dlogp, dHidden, dW.  `logits` (optional, [T,V] un-shifted): start from these logits instead of
    hidden @ W^T (same-logits checker, see `lm_head_backward`)."""
    logits = lm_head_logits(hidden, weight, round_bf16, dtype) if logits is None else np.asarray(logits, dtype=dtype)
    logp, ent, lse = logprobs_entropy(shift_rows(logits), ids, tau, dtype)
    lens = response_lengths(position_ids)
    loss, ratio, clipped, flag, dlogp = loss_fwd_bwd(logp, old, adv, mask, lens, cfg, loss_scale, dtype)
    d_hidden, d_weight = lm_head_backward(hidden, weight, ids, tau, lse, dlogp, round_bf16, dtype, logits=logits)
    return dict(
        logp=logp,
        entropy=ent,
        lse=lse,
        lengths=lens,
        loss=loss,
        importance_ratio=ratio,
        clipped_importance_ratio=clipped,
        is_clipped=flag,
        dlogp=dlogp,
        d_hidden=d_hidden,
        d_weight=d_weight,
    )


# --------------------------------------------------------------------------- #
# vocab-shard partials (online-softmax algebra; SURVEY §8a "single-formula restatement")
# --------------------------------------------------------------------------- #
def shard_partials(z_shard: np.ndarray, ids: np.ndarray, v0: int, dtype=np.float64) -> np.ndarray:
    """Per-token (m, s, u, z_target) of one vocab shard holding columns [v0, v0 + z_shard.shape[1]) of the
    temperature-scaled shifted logits z.  z_target is 0 when the target lies outside the shard."""
    z = z_shard.astype(dtype)
    T, Vs = z.shape
    if Vs == 0:
        out = np.zeros((T, 4), dtype=dtype)
        out[:, 0] = -np.inf
        return out
    m = z.max(axis=1)
    e = np.exp(z - m[:, None])
    s = e.sum(axis=1)
    u = (e * z).sum(axis=1)
    loc = np.asarray(ids).reshape(-1) - v0
    inside = (loc >= 0) & (loc < Vs)
    zt = np.where(inside, z[np.arange(T), np.clip(loc, 0, Vs - 1)], 0.0)
    return np.stack([m, s, u, zt], axis=1)


def merge_partials(parts: np.ndarray):
    """parts [n, T, 4] -> (logp, entropy, lse): rescale s,u by e^{m_i - max m}; lse = m + ln s; H = lse - u/s."""
    parts = np.asarray(parts, dtype=np.float64)
    m = parts[:, :, 0].max(axis=0)
    w = np.exp(parts[:, :, 0] - m[None, :])
    s = (parts[:, :, 1] * w).sum(axis=0)
    u = (parts[:, :, 2] * w).sum(axis=0)
    zt = parts[:, :, 3].sum(axis=0)
    lse = m + np.log(s)
    return zt - lse, lse - u / s, lse


# --------------------------------------------------------------------------- #
# SFT cross-entropy ("next" row N3: src/prime_rl/trainer/sft/train.py:209-216)
# --------------------------------------------------------------------------- #
def sft_cross_entropy(hidden, weight, targets, mask, round_bf16: bool, dtype=np.float64):
    """loss = mean over mask of -log_softmax(hidden @ W^T)[target] (targets pre-shifted, no temperature), and its
    gradients w.r.t. hidden and W.  Returns (loss, d_hidden, d_weight)."""
    logits = lm_head_logits(hidden, weight, round_bf16, dtype)
    t = np.asarray(targets).reshape(-1)
    mk = np.asarray(mask).reshape(-1).astype(bool)
    logp, _, lse = logprobs_entropy(logits, t, 1.0, dtype)
    n = max(int(mk.sum()), 1)
    loss = -(logp[mk]).sum() / n
    g = np.where(mk, -1.0 / n, 0.0)
    p = np.exp(logits - lse[:, None])
    dz = -p * g[:, None]
    dz[np.arange(len(t)), t] += g
    if round_bf16:
        dz = bf16_round(dz.astype(np.float32)).astype(dtype)
    return loss, dz @ weight.astype(dtype), dz.T @ hidden.astype(dtype)
