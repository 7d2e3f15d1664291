"""torch-CPU port of the reference chain, used ONLY as the timed CPU baseline
(`bench.py` `cpu_baseline` leg and `bench.py --impl reference`) and as a second
checker in tests.  TEST/BENCH INFRASTRUCTURE — never imported by the product.

It runs the same sequence of dense fp32 torch ops the reference trainer runs
between `hidden` and `loss.backward()` (src/prime_rl/trainer/rl/train.py:285-310),
restated from the formulas (not copied), so that the time it takes on the GPU
box's host cores is the reference's own CPU cost for the path.
"""

from __future__ import annotations

import torch

__all__ = ["segment_lengths", "advantages_cpu", "loss_cpu", "micro_step_cpu"]


def segment_lengths(position_ids: torch.Tensor) -> list[int]:
    """Vectorised equivalent of `get_response_lengths` (trainer/utils.py:36-69)."""
    p = position_ids.flatten()
    T = p.numel()
    if T < 3:
        return [T]
    i = torch.arange(1, T - 1, device=p.device)
    is_b = (p[1:-1] == 0) & (p[:-2] != 0) & (p[2:] == 1)
    b = [0] + i[is_b].tolist() + [T]
    return [b[k + 1] - b[k] for k in range(len(b) - 1)]


def advantages_cpu(rewards, lengths, group: int, kind: str) -> list[float]:
    """Per-group torch ops as in orchestrator/advantage.py:91-101."""
    out: list[float] = []
    for g0 in range(0, len(rewards), group):
        r = torch.tensor(rewards[g0 : g0 + group])
        if kind == "opo":
            w = torch.tensor(lengths[g0 : g0 + group]).to(r.dtype)
            a = r - (r * w).sum() / w.sum()
        else:
            a = r - r.mean()
            if kind == "drgrpo-negclipped":
                a = torch.maximum(a, torch.zeros_like(a))
            elif kind == "rloo":
                a = a * group / (group - 1)
        out.extend(a.tolist())
    return out


def loss_cpu(logp, old, adv, mask, lengths, ratio_type, length_norm, clip, loss_scale):
    """Segment loop of rl/loss.py:89-118 on 1-D tensors."""
    total = 0
    o = 0
    for n in lengths:
        lp, ol, a, mk = logp[o : o + n], old[o : o + n], adv[o : o + n], mask[o : o + n]
        o += n
        d = lp - ol
        if ratio_type == "sequence":
            s = d[mk].sum()
            if length_norm:
                s = s / torch.clamp_min(mk.sum(), 1)
            d = torch.clamp(s.unsqueeze(0), max=10.0)
        r = torch.exp(d)
        seg = (-torch.clamp(r, max=clip) * a)[mk].sum()
        if ratio_type == "sequence":
            seg = seg / torch.clamp_min(mk.sum(), 1)
        total = total + seg
    return total / max(loss_scale, 1)


def micro_step_cpu(hidden, weight, ids, position_ids, old, adv, mask, tau, ratio_type="token", length_norm=False,
                   clip=8.0, loss_scale=1):
    """One packed micro-batch: LM head -> fp32 logits -> shift -> /tau ->
    log-softmax gather -> loss -> entropy (no_grad) -> backward.  `hidden`
    [T,H] and `weight` [V,H] are leaf tensors; their `.grad` is filled.
    Returns (loss, logp, entropy)."""
    T = hidden.shape[0]
    logits = (hidden @ weight.t()).float().unsqueeze(0)                   # train.py:285
    shifted = torch.cat([torch.zeros_like(logits[:, :1]), logits[:, :-1]], dim=1)  # loss.py:54-57
    shifted = shifted / tau                                               # train.py:287
    sel = torch.gather(shifted, -1, ids.view(1, T, 1)).squeeze(-1)
    logp = (sel - torch.logsumexp(shifted, dim=-1)).squeeze(0)            # loss.py:26-30
    lens = segment_lengths(position_ids)
    loss = loss_cpu(logp, old, adv, mask, lens, ratio_type, length_norm, clip, loss_scale)
    with torch.no_grad():                                                 # train.py:302-303
        pd = torch.softmax(shifted, dim=-1)
        ent = (torch.logsumexp(shifted, dim=-1) - (pd * shifted).sum(-1)).squeeze(0)
    del logits, shifted
    loss.backward()                                                       # train.py:310
    return loss.detach(), logp.detach(), ent
