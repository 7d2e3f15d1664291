"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference functions (pmahdavi/prime-rl-cse) on seeded inputs.

Run in the build container only (the reference is not on the GPU box):

    PYTHONPATH=/root/reference/src python tests/golden/make_golden.py

Writes advantages.npz, segments.npz, loss.npz, lmhead.npz, sft.npz, packed_layout.npz next to this file.
The reference entry points exercised (paths relative to /root/reference):
  src/prime_rl/orchestrator/advantage.py:71   compute_advantages
  src/prime_rl/trainer/utils.py:36            get_response_lengths
  src/prime_rl/trainer/rl/loss.py:13,43,51,61 selective_log_softmax, compute_entropy,
                                              shift_logits, compute_loss
  src/prime_rl/trainer/rl/config.py:19        LossConfig
and the call-site glue of src/prime_rl/trainer/rl/train.py:285-310.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
if REF_SRC not in sys.path:
    sys.path.insert(0, REF_SRC)

from prime_rl.orchestrator.advantage import compute_advantages  # noqa: E402
from prime_rl.trainer.rl.config import LossConfig  # noqa: E402
from prime_rl.trainer.rl.loss import compute_entropy, compute_loss, selective_log_softmax, shift_logits  # noqa: E402
from prime_rl.trainer.utils import get_response_lengths  # noqa: E402

import prime_rl  # noqa: E402

assert prime_rl.__file__.startswith(REF_SRC), f"wrong prime_rl picked up: {prime_rl.__file__}"

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ["drgrpo", "drgrpo-negclipped", "rloo", "opo"]


def gen_advantages():
    rng = np.random.default_rng(1234)
    out = {}
    cases = []
    # KAT-A (docstring example, advantage.py:13-16,27-30)
    cases.append(("kat", [0.0, 0.0, 1.0, 1.0], [1, 2, 3, 4], 4))
    cases.append(("allzero", [0.0] * 4, [5, 6, 7, 8], 4))
    cases.append(("allone", [1.0] * 8, [3] * 8, 8))
    for G in [2, 3, 4, 8, 16, 24, 32]:
        n = G * 5
        cases.append((f"binary_G{G}", rng.integers(0, 2, n).astype(np.float32).tolist(), rng.integers(1, 4096, n).tolist(), G))
        cases.append((f"randn_G{G}", rng.standard_normal(n).astype(np.float32).tolist(), rng.integers(1, 4096, n).tolist(), G))
    # orders of reduction may differ from ATen's for these G: tolerance cases (<= 1e-6 rel)
    for G in [5, 12, 40, 64, 100]:
        n = G * 3
        cases.append((f"tol_randn_G{G}", rng.standard_normal(n).astype(np.float32).tolist(), rng.integers(1, 4096, n).tolist(), G))
    # config 3 / config 5 shapes
    cases.append(("cfg3", rng.integers(0, 2, 1024).astype(np.float32).tolist(), rng.integers(16, 4096, 1024).tolist(), 16))
    cases.append(("cfg5", rng.random(32 * 4).astype(np.float32).tolist(), rng.integers(1024, 32768, 128).tolist(), 32))
    cases.append(("G1", [0.25, 0.5], [3, 4], 1))  # rloo -> nan, others 0
    names = []
    for name, r, ln, G in cases:
        names.append(name)
        out[f"{name}/rewards"] = np.asarray(r, dtype=np.float32)
        out[f"{name}/lengths"] = np.asarray(ln, dtype=np.int64)
        out[f"{name}/group"] = np.asarray(G)
        for k in KINDS:
            a = compute_advantages(r, ln, G, k)
            out[f"{name}/{k}"] = np.asarray(a, dtype=np.float32)
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "advantages.npz"), **out)


def gen_segments():
    rng = np.random.default_rng(7)
    cases = {
        "kat2a": [0, 1, 2, 0, 1, 2, 3, 0, 1],
        "kat2b": [0, 1, 2, 0, 1, 0, 0, 0],
        "kat2c": [0, 1, 2, 0, 0, 1, 2],
        "kat2d": list(range(4)) * 3,
        "allzero": [0] * 16,
        "single": [0],
        "two": [0, 1],
        "three_restart": [0, 0, 1],
        "len1_segments": [0, 0, 0, 1, 0, 1, 2],
        "trailing_pad": [0, 1, 2, 3, 0, 1, 0, 0],
    }
    for i in range(6):
        lens = rng.integers(1, 40, rng.integers(1, 9))
        pos = np.concatenate([np.arange(n) for n in lens])
        if i % 2:
            pos = np.concatenate([pos, np.zeros(rng.integers(0, 5), dtype=pos.dtype)])
        cases[f"rand{i}"] = pos.tolist()
    out, names = {}, []
    for name, pos in cases.items():
        names.append(name)
        out[f"{name}/position_ids"] = np.asarray(pos, dtype=np.int64)
        out[f"{name}/lengths"] = np.asarray(get_response_lengths(torch.tensor(pos)), dtype=np.int64)
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "segments.npz"), **out)


def run_ref_loss(logp, old, adv, mask, lens, cfg, scale):
    lp = torch.tensor(logp, requires_grad=True)
    loss, t = compute_loss(
        lp.split(lens),
        torch.tensor(old).split(lens),
        torch.tensor(adv).split(lens),
        torch.tensor(mask).split(lens),
        cfg,
        scale,
    )
    loss.backward()
    return (
        loss.detach().numpy(),
        t["importance_ratio"].detach().numpy(),
        t["clipped_importance_ratio"].detach().numpy(),
        t["is_clipped"].detach().numpy(),
        lp.grad.numpy(),
    )


def gen_loss():
    rng = np.random.default_rng(99)
    out, names = {}, []
    base = []
    # KAT-3 / KAT-4 data
    base.append(("kat", [-1, -2, -0.5, -3, -0.1], [-1.5, -1, -3, -3, -0.2], [0.5, 0.5, 0.5, -1, -1], [1, 1, 1, 0, 1], [3, 2], None))
    for i, (lens, sigma, pmask) in enumerate(
        [([50, 30], 0.15, 1.0), ([40, 60], 1.5, 0.7), ([7, 1, 128, 33, 2], 0.5, 0.5), ([200], 3.0, 0.9), ([16, 16, 16], 0.1, 0.0)]
    ):
        T = sum(lens)
        logp = -np.abs(rng.standard_normal(T)).astype(np.float32)
        old = (logp + sigma * rng.standard_normal(T)).astype(np.float32)
        adv = np.concatenate([np.full(n, rng.standard_normal(), dtype=np.float32) for n in lens])
        mask = rng.random(T) < pmask
        base.append((f"rand{i}", logp, old, adv, mask, lens, None))
    cfgs = {
        "token": LossConfig(ratio_type="token", clip_ratio=8.0),
        "token_c2": LossConfig(ratio_type="token", clip_ratio=2.0),
        "seq": LossConfig(ratio_type="sequence", ratio_length_norm=False, clip_ratio=8.0),
        "seqnorm": LossConfig(ratio_type="sequence", ratio_length_norm=True, clip_ratio=8.0),
    }
    for name, logp, old, adv, mask, lens, _ in base:
        logp = np.asarray(logp, dtype=np.float32)
        old = np.asarray(old, dtype=np.float32)
        adv = np.asarray(adv, dtype=np.float32)
        mask = np.asarray(mask).astype(bool)
        names.append(name)
        out[f"{name}/logp"], out[f"{name}/old"], out[f"{name}/adv"], out[f"{name}/mask"] = logp, old, adv, mask
        out[f"{name}/lens"] = np.asarray(lens, dtype=np.int64)
        for cname, cfg in cfgs.items():
            for scale in ([4, 2, 0] if name == "kat" else [max(int(mask.sum()), 0), len(lens)]):
                res = run_ref_loss(logp, old, adv, mask, list(lens), cfg, scale)
                for key, v in zip(["loss", "ratio", "clipped", "is_clipped", "dlogp"], res):
                    out[f"{name}/{cname}/S{scale}/{key}"] = np.asarray(v)
    out["names"] = np.asarray(names)
    out["cfgs"] = np.asarray(list(cfgs.keys()))
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


def ref_chain(hidden, weight, ids, pos, old, adv, mask, tau, cfg, scale):
    """train.py:285-310 with the reference functions; `hidden`/`weight` are leaf
    tensors in the model dtype (fp32 for config 1, bf16 for the GPU configs)."""
    h = hidden.clone().requires_grad_(True)
    w = weight.clone().requires_grad_(True)
    logits = (h @ w.t()).unsqueeze(0).float().contiguous()
    shifted = shift_logits(logits) / tau
    logprobs = selective_log_softmax(shifted, ids.unsqueeze(0))
    lens = get_response_lengths(pos)
    loss, tens = compute_loss(
        logprobs.squeeze().split(lens), old.split(lens), adv.split(lens), mask.split(lens), cfg, scale
    )
    with torch.no_grad():
        ent = compute_entropy(shifted)
    loss.backward()
    return dict(
        logp=logprobs.detach().squeeze(0).numpy(),
        entropy=ent.squeeze(0).numpy(),
        loss=loss.detach().numpy(),
        ratio=tens["importance_ratio"].detach().numpy(),
        d_hidden=h.grad.float().numpy(),
        d_weight=w.grad.float().numpy(),
        lens=np.asarray(lens, dtype=np.int64),
    )


def gen_lmhead():
    g = torch.Generator().manual_seed(4321)
    out, names = {}, []
    # KAT-1: identity "LM head" so logits are given directly
    logits = torch.tensor([[1.0, 2, 3, 4], [0, 0, 0, 0], [4, 3, 2, 1]])
    ids = torch.tensor([2, 3, 0])
    for tau in (1.0, 2.0):
        sh = shift_logits(logits.unsqueeze(0)) / tau
        out[f"kat1/tau{tau}/logp"] = selective_log_softmax(sh, ids.unsqueeze(0)).squeeze(0).numpy()
        out[f"kat1/tau{tau}/entropy"] = compute_entropy(sh).squeeze(0).numpy()
    out["kat1/logits"], out["kat1/ids"] = logits.numpy(), ids.numpy()

    shapes = [("small_fp32", 48, 32, 160, torch.float32), ("small_bf16", 48, 32, 160, torch.bfloat16),
              ("mid_bf16", 160, 64, 1000, torch.bfloat16)]
    for name, T, H, V, dt in shapes:
        hidden = torch.randn(T, H, generator=g).to(dt)
        weight = (torch.randn(V, H, generator=g) * (4.0 / H**0.5)).to(dt)
        lens = [T // 3, T // 4, T - T // 3 - T // 4]
        pos = torch.cat([torch.arange(n) for n in lens])
        ids = torch.randint(0, V, (T,), generator=g)
        adv = torch.cat([torch.full((n,), float(torch.randn((), generator=g))) for n in lens])
        mask = torch.rand(T, generator=g) < 0.75
        for tau in (1.0, 0.6):
            with torch.no_grad():
                lg = (hidden @ weight.t()).float().unsqueeze(0)
                lp_true = selective_log_softmax(shift_logits(lg) / tau, ids.unsqueeze(0)).squeeze(0)
            old = lp_true + 0.3 * torch.randn(T, generator=g)
            for cname, cfg, scale in [
                ("token", LossConfig(), int(mask.sum())),
                ("seqnorm", LossConfig(ratio_type="sequence", ratio_length_norm=True), 1),
            ]:
                res = ref_chain(hidden, weight, ids, pos, old, adv, mask, tau, cfg, scale)
                key = f"{name}/tau{tau}/{cname}"
                names.append(key)
                for k, v in res.items():
                    out[f"{key}/{k}"] = v
                out[f"{key}/old"] = old.numpy()
                out[f"{key}/scale"] = np.asarray(scale)
        out[f"{name}/hidden"] = hidden.float().numpy()
        out[f"{name}/weight"] = weight.float().numpy()
        out[f"{name}/ids"], out[f"{name}/pos"] = ids.numpy(), pos.numpy()
        out[f"{name}/adv"], out[f"{name}/mask"] = adv.numpy(), mask.numpy()
        out[f"{name}/is_bf16"] = np.asarray(dt == torch.bfloat16)
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "lmhead.npz"), **out)


def gen_sft():
    """SFT cross-entropy call site (src/prime_rl/trainer/sft/train.py:209-216), restated with the same torch call."""
    from torch.nn.functional import cross_entropy

    g = torch.Generator().manual_seed(77)
    out = {}
    for name, T, H, V, dt in [("sft_fp32", 64, 32, 200, torch.float32), ("sft_bf16", 96, 64, 520, torch.bfloat16)]:
        hidden = torch.randn(1, T, H, generator=g).to(dt).requires_grad_(True)
        weight = (torch.randn(V, H, generator=g) * (4.0 / H**0.5)).to(dt).requires_grad_(True)
        targets = torch.randint(0, V, (1, T), generator=g)
        mask = torch.rand(1, T, generator=g) < 0.7
        logits = hidden @ weight.t()
        B, L, Vv = logits.shape
        loss = cross_entropy(logits.view(-1, Vv), targets.view(-1), reduction="none").view(B, L)
        loss = loss[mask].mean()
        loss.backward()
        out[f"{name}/hidden"], out[f"{name}/weight"] = hidden.detach().float().numpy(), weight.detach().float().numpy()
        out[f"{name}/targets"], out[f"{name}/mask"] = targets.numpy(), mask.numpy()
        out[f"{name}/loss"] = loss.detach().float().numpy()
        out[f"{name}/d_hidden"], out[f"{name}/d_weight"] = hidden.grad.float().numpy(), weight.grad.float().numpy()
        out[f"{name}/is_bf16"] = np.asarray(dt == torch.bfloat16)
    out["names"] = np.asarray(["sft_fp32", "sft_bf16"])
    np.savez_compressed(os.path.join(HERE, "sft.npz"), **out)


def gen_packed_layout():
    """KAT-5: the packed MicroBatch layout exactly as the reference's orchestrator produces it
    (prepare_batch, src/prime_rl/orchestrator/batch.py:123-175) from toy rollouts whose advantages come from the
    reference's compute_advantages: prompt tokens masked, multi-turn holes in the completion mask, per-rollout advantage
    broadcast over prompt + completion, first-fit-decreasing packing, all-masked padding micro-batches for rank balance."""
    from prime_rl.orchestrator.batch import prepare_batch
    from prime_rl.orchestrator.buffer import Rollout

    rng = np.random.default_rng(2024)
    V, G, n_problems = 1000, 4, 3
    rewards, lens, protos = [], [], []
    for p in range(n_problems):
        prompt = rng.integers(0, V, rng.integers(3, 9)).tolist()
        for _ in range(G):
            n = int(rng.integers(4, 40))
            comp = rng.integers(0, V, n).tolist()
            cmask = (rng.random(n) < 0.8).astype(int).tolist()  # environment turns masked out
            clogp = (-rng.random(n) * 3).astype(np.float32).tolist()
            r = float(rng.integers(0, 2))
            rewards.append(r)
            lens.append(n)
            protos.append((p, prompt, comp, cmask, clogp, r))
    adv = compute_advantages(rewards, lens, G, "drgrpo")
    rollouts = [Rollout(problem_id=p, prompt_tokens=pr, prompt_mask=[0] * len(pr), completion_tokens=c, completion_mask=cm,
                        completion_logprobs=cl, is_truncated=False, reward=r, advantage=a)
                for (p, pr, c, cm, cl, r), a in zip(protos, adv)]
    per_rank = prepare_batch(rollouts, temperature=0.6, tokenizer=None, batch_size=len(rollouts), micro_batch_size=2, seq_len=48,
                             num_train_workers=2)
    out = {"n_ranks": np.asarray(len(per_rank)), "rewards": np.asarray(rewards, dtype=np.float32), "lens": np.asarray(lens),
           "advantages": np.asarray(adv, dtype=np.float32)}
    # the rollouts themselves (inputs of prepare_batch), flattened
    out["ro/prompt_len"] = np.asarray([len(r.prompt_tokens) for r in rollouts], dtype=np.int64)
    out["ro/completion_len"] = np.asarray([len(r.completion_tokens) for r in rollouts], dtype=np.int64)
    out["ro/prompt_tokens"] = np.concatenate([np.asarray(r.prompt_tokens, dtype=np.int64) for r in rollouts])
    out["ro/prompt_mask"] = np.concatenate([np.asarray(r.prompt_mask, dtype=np.int64) for r in rollouts])
    out["ro/completion_tokens"] = np.concatenate([np.asarray(r.completion_tokens, dtype=np.int64) for r in rollouts])
    out["ro/completion_mask"] = np.concatenate([np.asarray(r.completion_mask, dtype=np.int64) for r in rollouts])
    out["ro/completion_logprobs"] = np.concatenate([np.asarray(r.completion_logprobs, dtype=np.float32) for r in rollouts])
    out["ro/advantage"] = np.asarray([r.advantage for r in rollouts], dtype=np.float32)
    out["ro/args"] = np.asarray([0.6, len(rollouts), 2, 48, 2])  # temperature, batch_size, micro_batch_size, seq_len, num_train_workers
    for r, mbs in enumerate(per_rank):
        out[f"rank{r}/n"] = np.asarray(len(mbs))
        for i, mb in enumerate(mbs):
            for k in ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask"):
                out[f"rank{r}/mb{i}/{k}"] = mb[k].numpy()
            out[f"rank{r}/mb{i}/temperature"] = np.asarray(mb["temperature"])
            out[f"rank{r}/mb{i}/lengths"] = np.asarray(get_response_lengths(mb["position_ids"]), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "packed_layout.npz"), **out)
    gen_rollout_files(per_rank)


def gen_rollout_files(per_rank, step=7):
    """The orchestrator -> trainer on-disk format: `<output_dir>/rollouts/step_<n>/rank_<r>.pt` = torch.save(list[MicroBatch]),
    written to a .tmp file and renamed (src/prime_rl/orchestrator/orchestrator.py:356-363), read back by the reference's own
    `DataLoader` (src/prime_rl/trainer/rl/data.py:57-75) to confirm the fixture is what the trainer consumes."""
    from pathlib import Path

    from prime_rl.utils.utils import get_rollout_dir

    step_path = get_rollout_dir(Path(HERE)) / f"step_{step}"
    step_path.mkdir(parents=True, exist_ok=True)
    for i, batches in enumerate(per_rank):
        batch_path = step_path / f"rank_{i}.pt"
        tmp_path = batch_path.with_suffix(".tmp")
        torch.save(batches, tmp_path)
        tmp_path.rename(batch_path)
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("LOCAL_WORLD_SIZE", "1")
    from prime_rl.trainer.rl.data import DataLoader

    import logging

    import prime_rl.utils.logger as ref_logger

    ref_logger._LOGGER = logging.getLogger("golden")  # wait_for_path logs through the reference's global logger
    dl = DataLoader(Path(HERE), step)
    dl.wait_for_batch()
    got = dl.get_batch()
    assert len(got) == len(per_rank[0]) and all(torch.equal(a["input_ids"], b["input_ids"]) for a, b in zip(got, per_rank[0]))
    print(f"rollouts/step_{step}: {len(per_rank)} rank files, read back by the reference DataLoader")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for reproducible fixtures
    gens = dict(advantages=gen_advantages, segments=gen_segments, loss=gen_loss, lmhead=gen_lmhead, sft=gen_sft,
                packed_layout=gen_packed_layout)
    for name in (sys.argv[1:] or list(gens)):  # `python make_golden.py packed_layout` regenerates one family
        gens[name]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
