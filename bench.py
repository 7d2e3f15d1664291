#!/usr/bin/env python
"""bench.py — PRIME-RL per-token RL loss hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one rank's share of a training batch at BASELINE.json's Qwen3-8B shape
(configs[2], the config the metric is quoted on): 32 packed micro-batches of [1, 16384] tokens (4 x 4096-token rollouts
each), hidden 4096, vocab 151936, bf16 -- per micro-batch: fused LM-head forward (log-probs + entropy), packed clipped-ratio
loss, backward to d_hidden and d_W_lm (fp32 accumulation across micro-batches), plus the step's group advantages.  Weak
scaling: every rank processes its own 32 micro-batches (token shards = FSDP ranks); for N > 1 the step also delivers d_W_lm in
FSDP2's reduce-scattered layout -- by the last micro-batch's d_weight GEMM epilogue over NVLink peer memory in the recompute mode
(`--fused-dw-sync`), by one ncclReduceScatter per step in the saved-logits mode.

ONE run times BOTH backward modes of the LM head:
  * recompute (the HEADLINE: top-level `value`, `e2e`, `roofline`): BASELINE.json north_star's design -- the [T, V] logits
    never exist in HBM, the backward recomputes them per vocab chunk (4 GEMMs of T*H*V per micro-batch, 3 credited);
  * saved_logits (`modes.saved_logits`, a labelled opt-in): one bf16 [T, V] block per micro-batch is kept between forward
    and backward (3 GEMMs, 4.6 GiB).
`regimes` holds short end-to-end runs of the other measurement regimes of SURVEY §8(d) (peaked logits, tau = 0.6, a clipping-
heavy run, prompt and multi-turn masks, sequence-level loss on 32k-token rows, the 1.7B and 0.6B shapes).
`--workload qwen3-32b-vocab8` (torchrun x N) is BASELINE config 4: the LM head vocab-sharded over the ranks.

Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# stdout carries exactly ONE JSON line: everything else that libraries print to fd 1 (e.g. NCCL's version banner) is sent to
# stderr by pointing fd 1 at fd 2 for the duration of the run; emit() writes the result to the saved, real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


METRIC = "loss_fwd_bwd_tokens_per_sec"
UNIT = "tokens/s"

SHAPES = {
    # name: (H, V, tokens per micro-batch, segments per micro-batch, micro-batches per rank per step, group size, ratio_type)
    "qwen3-8b-hendrycks": (4096, 151936, 16384, 4, 32, 16, "token"),
    "qwen3-1.7b-wordle": (2048, 151936, 16384, 8, 8, 16, "token"),
    "qwen3-0.6b-reverse": (1024, 151936, 4096, 32, 1, 8, "token"),
    "qwen3-8b-deepscaler": (4096, 151936, 32768, 1, 8, 32, "sequence"),
    "qwen3-32b-vocab8": (5120, 151936, 16384, 4, 8, 16, "token"),  # BASELINE config 4: run under torchrun, vocab-sharded
    "tiny": (128, 2048, 512, 4, 2, 4, "token"),  # CI smoke of this script only, not a benchmark
}
VCHUNK = 16384  # default vocab chunk of prl_lmhead_bwd


# algorithmic bytes per launch (operands read once + outputs written once), for comparison with `traffic`; vc = vocab columns
# one launch covers (the vocab chunk in recompute mode, V in saved-logits mode)
def algorithmic_bytes(kernel, T, H, V, vc):
    if kernel == "lmhead_fwd_kernel":
        return 2.0 * (T * H + V * H) + 16.0 * T
    if kernel == "lmhead_fwd_save_kernel":
        return 2.0 * (T * H + V * H) + 16.0 * T + 2.0 * T * V
    if kernel == "dz_from_saved_logits_kernel":
        return 4.0 * T * V + 16.0 * T
    if kernel == "lmhead_bwd_dz_kernel":
        return 2.0 * (T * H + vc * H) + 2.0 * T * vc
    if kernel == "lmhead_bwd_dhidden_kernel":
        return 2.0 * (T * vc + vc * H) + (8.0 if vc < V else 2.0) * T * H
    return 2.0 * (T * vc + T * H) + 8.0 * vc * H


def ncu_traffic(mode: str, kernel: str, workload: str):
    """DRAM read+write bytes per launch of `kernel` from the newest committed ncu capture (profiles/*traffic.json, written by
    tools/make_evidence.py from `ncu --set full` reports): {"bytes": ..., "source": ...} or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for name in sorted(os.listdir(pdir)):
            if name.endswith("traffic.json"):
                try:
                    with open(os.path.join(pdir, name)) as f:
                        d = json.load(f)
                    rec = d.get(workload, {}).get(mode, {}).get(kernel)
                    if rec:
                        best = {"bytes": rec["dram_bytes_per_launch"], "source": f"profiles/{name}: {rec.get('source', '')}"}
                except Exception:  # noqa: BLE001
                    continue
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clock sampling during the timed region (B200_PROFILING.md "clocks" line)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        self.index = index

    def start(self):
        self.rows = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.proc = None
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
DEFAULT_REGIME = dict(dist="uniform", tau=1.0, sigma=0.15, prompt_frac=0.0, multi_turn=False)


def make_workload(shape: str, rank: int, n_distinct: int, pin: bool, device="cpu", regime: dict | None = None, n_mb: int = 0):
    """Micro-batches of the named shape, generated on `device` (the GPU for the B200 arm: seconds instead of a single-threaded
    CPU randn under torchrun) and returned as HOST tensors (pinned when `pin`), exactly what a data loader hands the trainer;
    `weight` stays on `device`.  Regime: dist "uniform" = W ~ N(0, 0.02^2), hidden ~ N(0, 1), ids uniform
    (near-uniform softmax, worst case: nothing is sparse); "peaked" = a rank-1 answer component lifts the sampled token's logit
    by N(ln V + 3, 2^2) (mean entropy ~2-3 nats instead of ln V = 11.9: closer to what RL sees).  `prompt_frac` masks the first part of every rollout
    (mask 0, old log-prob 0: orchestrator/batch.py:41-42), `multi_turn` additionally masks alternating environment runs.
    Old log-probs here are a placeholder near ln(1/V); the B200 arm replaces them by (its own log-probs + N(0, sigma^2))."""
    rg = dict(DEFAULT_REGIME, **(regime or {}))
    H, V, T, nseg, n_mb_shape, group, ratio_type = SHAPES[shape]
    n_mb = n_mb if n_mb > 0 else n_mb_shape
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    weight = (torch.randn(V, H, generator=g, dtype=torch.float32, device=device) * 0.02).to(torch.bfloat16)
    seg = [T // nseg] * nseg
    seg[-1] += T - sum(seg)
    mbs = []
    n_rollouts = nseg * n_mb
    n_rollouts += (-n_rollouts) % group
    rewards = (torch.rand(n_rollouts, generator=g, device=device) < 0.5).float().cpu()
    lengths = torch.full((n_rollouts,), T // nseg, dtype=torch.int64)
    adv_r = (rewards.view(-1, group) - rewards.view(-1, group).mean(dim=1, keepdim=True)).reshape(-1)  # drgrpo, data only
    for i in range(n_distinct):
        hidden = torch.randn(T, H, generator=g, dtype=torch.float32, device=device)
        ids = torch.randint(0, V, (T,), generator=g, device=device)
        if rg["dist"] == "peaked":
            tgt = weight[ids[1:]].float()  # position t is predicted from hidden row t-1 (shift_logits)
            lift = math.log(V) + 3.0 + 2.0 * torch.randn(T - 1, 1, generator=g, device=device)
            hidden[:-1] += lift * tgt / (tgt * tgt).sum(-1, keepdim=True)
        hidden = hidden.to(torch.bfloat16).cpu()
        ids = ids.cpu()
        pos = torch.cat([torch.arange(n) for n in seg])
        mask = torch.ones(T, dtype=torch.bool)  # headline: all ones, as FakeDataLoader (rl/data.py:53)
        o = 0
        for n in seg:
            p = int(n * rg["prompt_frac"])
            mask[o : o + p] = False
            if rg["multi_turn"]:  # wordle-style: model turn (mask 1) / environment turn (mask 0), orchestrator/config.py:463-468
                turn = max(8, n // 16)
                k = o + p + turn
                while k < o + n:
                    mask[k : min(k + turn // 2, o + n)] = False
                    k += 2 * turn
            o += n
        old = (-11.9 + rg["sigma"] * torch.randn(T, generator=g, device=device)).cpu()
        old[~mask] = 0.0
        mb = dict(hidden=hidden.unsqueeze(0), input_ids=ids.unsqueeze(0), position_ids=pos.unsqueeze(0),
                  advantages=torch.cat([adv_r[(i * nseg + j) % n_rollouts].repeat(n) for j, n in enumerate(seg)]).unsqueeze(0),
                  logprobs=old.unsqueeze(0), loss_mask=mask.unsqueeze(0), temperature=float(rg["tau"]))
        if pin:
            for k, v in mb.items():
                if torch.is_tensor(v):
                    mb[k] = v.pin_memory()
        mbs.append(mb)
    return dict(shape=shape, H=H, V=V, T=T, seg=seg, n_mb=n_mb, group=group, ratio_type=ratio_type, weight=weight, mbs=mbs,
                rewards=rewards, lengths=lengths, regime=rg)


def describe(wl, world=1) -> str:
    rg = wl["regime"]
    mask = ("all-ones mask" if rg["prompt_frac"] == 0 and not rg["multi_turn"] else
            f"first {rg['prompt_frac']:.0%} of every rollout masked (prompt)" + (" + alternating environment turns masked" if rg["multi_turn"] else ""))
    return (f"{wl['shape']}: H={wl['H']} V={wl['V']}, {wl['n_mb']} packed micro-batches x [1,{wl['T']}] tokens per GPU per step "
            f"({len(wl['seg'])} x {wl['seg'][0]}-token rollouts each), ratio_type={wl['ratio_type']}, {mask}, tau={rg['tau']}, "
            f"D-{rg['dist']}, old log-probs = new + N(0, {rg['sigma']}^2)")


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
class Runner:
    """One workload resident on one GPU: the raw C-ABI loop (`step_resident`) and the public-API loop from host buffers
    (`step_e2e`), in either backward mode."""

    def __init__(self, wl, dev, dist, world, args):
        import prime_rl_cse_b200 as P
        from prime_rl_cse_b200 import _lib

        self.P, self._lib, self.wl, self.dev, self.dist, self.world, self.args = P, _lib, wl, dev, dist, world, args
        H, V, T = wl["H"], wl["V"], wl["T"]
        self.n_mb = wl["n_mb"]
        self.cfg = P.LossConfig(ratio_type=wl["ratio_type"], ratio_length_norm=wl["ratio_type"] == "sequence", clip_ratio=8.0)
        self.mode_id = 0 if self.cfg.ratio_type == "token" else 1
        self.W = wl["weight"].to(dev).requires_grad_(True)  # lm_head.weight is a trainable parameter
        self.Wd = self.W.detach()
        self.main_grad = torch.zeros(V, H, dtype=torch.float32, device=dev)  # fp32 accumulation across micro-batches
        self.rs_out = torch.empty((V // world) * H, dtype=torch.float32, device=dev) if world > 1 and V % world == 0 else None
        lib = _lib.load()
        self.ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
        self.ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 0), dtype=torch.uint8, device=dev)
        self.ws_bs = torch.empty(max(int(lib.prl_lmhead_bwd_saved_workspace_bytes(T, H, V)), 16), dtype=torch.uint8, device=dev)
        self.saved_blk = None
        # old log-probs = the path's own log-probs + N(0, sigma^2) (0 at masked positions), so that ratios are ~ exp(N(0, sigma^2))
        g = torch.Generator(device=dev).manual_seed(99)
        for mb in wl["mbs"]:
            lp, _, _, _ = P.lm_head_forward_raw(mb["hidden"][0].to(dev), self.Wd, mb["input_ids"][0].to(dev), 1.0 / mb["temperature"],
                                                workspace=self.ws_f)
            noise = wl["regime"]["sigma"] * torch.randn(T, generator=g, device=dev)
            if self.cfg.ratio_type == "sequence":
                noise = noise / math.sqrt(T / len(wl["seg"]))
            old = (lp + noise) * mb["loss_mask"][0].to(dev)
            mb["logprobs"].copy_(old.unsqueeze(0).cpu())
        self.dev_mbs = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in mb.items()} for mb in wl["mbs"]]
        self.rewards_d, self.lengths_d = wl["rewards"].to(dev), wl["lengths"].to(dev)
        bounds = [0]
        for n in wl["seg"]:
            bounds.append(bounds[-1] + n)
        self.cu = torch.tensor(bounds, dtype=torch.int32, device=dev)
        n_unmasked = int(wl["mbs"][0]["loss_mask"].sum())
        self.loss_scale = self.n_mb * n_unmasked if self.mode_id == 0 else self.n_mb  # rl/train.py:263-267
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in wl["mbs"][0].values() if torch.is_tensor(v)) * self.n_mb
        self.loss_host = torch.empty(self.n_mb, dtype=torch.float32).pin_memory()
        self.rs_events = None  # (start, end) CUDA events around the last step's d_W_lm reduce-scatter
        self.compute_end = None
        # N > 1, recompute mode: d_W_lm is accumulated straight into the reduce-scattered layout by the d_weight GEMM's epilogue
        # (peer memory over NVLink; prime_rl_cse_b200.peer.ShardedGradAccumulator) -- no step-end collective.  Falls back to the
        # local fp32 accumulator + ncclReduceScatter when peer buffers cannot be had (all ranks agree on that together).
        self.acc, self.acc_note = None, "local fp32 accumulator + ncclReduceScatter once per step"
        if dist is not None and world <= 8 and getattr(args, "fused_dw_sync", 1):
            try:
                every = int(args.fused_dw_sync) == 2
                self.acc = P.ShardedGradAccumulator(V, H, dev, every_micro_batch=every, local=None if every else self.main_grad)
                self.acc_note = ("every micro-batch's d_weight GEMM epilogue adds its tiles straight into the owner's shard (TMA reduce-add over NVLink; no "
                                 "rank holds a [V, H] gradient), barrier at step end" if every else
                                 "local fp32 accumulation; the LAST micro-batch's d_weight GEMM epilogue sends (local total + its tiles) straight into the "
                                 "owner's shard (TMA reduce-add over NVLink, hidden behind that micro-batch's GEMMs), barrier at step end")
            except RuntimeError as e:
                self.acc_note += f" (fused accumulation unavailable: {e})"[:200]

    def ensure_saved_block(self):
        if self.saved_blk is None:
            lib = self._lib.load()
            self.saved_blk = torch.empty(self.wl["T"], int(lib.prl_lmhead_saved_logits_ld(self.wl["V"])), dtype=torch.bfloat16, device=self.dev)

    def _grad_sync(self, fused: bool = False):
        """FSDP2's gradient sync of the [lm_head, norm] group (model.py:126-131): fp32 reduce-scatter of d_W_lm, once per step;
        `fused`: the shards were accumulated in place by the GEMM epilogues, what is left is the barrier"""
        if fused:
            self.compute_end = torch.cuda.Event(enable_timing=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.compute_end.record()
            a.record()
            self.acc.finish_step()
            b.record()
            self.rs_events = (a, b)
            return
        if self.dist is not None and self.rs_out is not None:
            self.compute_end = torch.cuda.Event(enable_timing=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.compute_end.record()
            a.record()
            self.dist.reduce_scatter_tensor(self.rs_out, self.main_grad.view(-1))
            b.record()
            self.rs_events = (a, b)

    def step_resident(self, save: bool):
        """inputs already in HBM; raw C-ABI calls, no autograd bookkeeping"""
        P, wl = self.P, self.wl
        P.compute_advantages_tensor(self.rewards_d, self.lengths_d, wl["group"], "drgrpo")
        last = None
        fused = self.acc is not None and not save
        if fused:
            self.acc.begin_step(self.n_mb)
        if save:
            self.ensure_saved_block()
        for i in range(self.n_mb):
            mb = self.dev_mbs[i % len(self.dev_mbs)]
            hid, ids, inv = mb["hidden"][0], mb["input_ids"][0], 1.0 / mb["temperature"]
            if save:
                logp, ent, lse, _, _ = P.lm_head_forward_save_raw(hid, self.Wd, ids, inv, workspace=self.ws_f, saved=self.saved_blk)
            else:
                logp, ent, lse, _ = P.lm_head_forward_raw(hid, self.Wd, ids, inv, workspace=self.ws_f)
            from prime_rl_cse_b200.loss import _loss_packed_raw

            loss, ratio, clipped, is_clipped, dlogp = _loss_packed_raw(
                logp, mb["logprobs"][0], mb["advantages"][0], mb["loss_mask"][0].view(torch.uint8), self.cu, len(wl["seg"]), self.mode_id,
                self.cfg.ratio_length_norm, self.cfg.clip_ratio, self.loss_scale)
            if save:
                P.lm_head_backward_saved_raw(hid, self.Wd, ids, lse, dlogp, inv, self.saved_blk, dweight=self.main_grad,
                                             dweight_accumulate=True, workspace=self.ws_bs)
            elif fused:
                self.acc.accumulate(hid, self.Wd, ids, lse, dlogp, inv, workspace=self.ws_b)
            else:
                P.lm_head_backward_raw(hid, self.Wd, ids, lse, dlogp, inv, dweight=self.main_grad, dweight_accumulate=True,
                                       workspace=self.ws_b)
            last = loss
        self._grad_sync(fused)
        return last

    def step_e2e(self, save: bool, forward_skips_masked: bool = False):
        """public API (`rl_micro_step`), HOST micro-batches: H2D of every micro-batch from pinned memory
        (prefetched one ahead on a copy stream), D2H of every micro-batch's loss."""
        P, wl, dev = self.P, self.wl, self.dev
        adv = P.compute_advantages(wl["rewards"].tolist(), wl["lengths"].tolist(), wl["group"], "drgrpo")  # host lists in/out
        del adv
        cur = torch.cuda.current_stream()
        fused = self.acc is not None and not save
        if fused:
            self.acc.begin_step(self.n_mb)

        def upload(i):
            src = wl["mbs"][i % len(wl["mbs"])]
            with torch.cuda.stream(self.copy_stream):
                out = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in src.items()}
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            # what a loader knows from the host copy (rollout_io.upload_micro_batches does the same): the micro-step then needs
            # no device-to-host synchronisation for the segment lengths (train.py:291) or the active-position count
            out["response_lengths"] = P.response_lengths_host(src["position_ids"])
            out["n_active"] = int(src["loss_mask"].reshape(-1)[1:].sum())
            return out, ev

        nxt = upload(0)
        for i in range(self.n_mb):
            mb, ev = nxt
            if i + 1 < self.n_mb:
                nxt = upload(i + 1)
            cur.wait_event(ev)
            for v in mb.values():
                if torch.is_tensor(v):
                    v.record_stream(cur)
            hid = mb["hidden"].requires_grad_(True)
            out = P.rl_micro_step(hid, self.W, mb, self.cfg, self.loss_scale, main_grad=self.acc if fused else self.main_grad,
                                  response_lengths=None, forward_skips_masked=forward_skips_masked, save_logits=save)
            self.loss_host[i].copy_(out["loss"], non_blocking=True)
        self._grad_sync(fused)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host[-1])

    # ---- timing ----
    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed warm-up steps, then K steps bracketed by barrier + synchronize; CUDA events; max over ranks.
        Returns (ms for K steps [max over ranks], this rank's ms, this rank's d_W reduce-scatter ms of the last step or None)."""
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        rs_ms = self.rs_events[0].elapsed_time(self.rs_events[1]) if self.rs_events is not None else None
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), ms, rs_ms

    def dw_sync_check(self):
        """the fused accumulation against the classic path on the same micro-batches: one step each from zeroed accumulators;
        every rank compares its shard (sum over ranks) -- rel-L2 and the largest element-wise deviation relative to the largest
        element (the ranks' addends reach a shard in no fixed order: fp32 rounding, not bits)"""
        if self.acc is None or self.rs_out is None:
            return None
        self.main_grad.zero_()
        self.step_resident_classic()
        torch.cuda.synchronize()
        ref = self.rs_out.view(-1, self.wl["H"]).clone()
        self.step_resident(False)
        torch.cuda.synchronize()
        got = self.acc.shard
        vals = [(got - ref).norm() / ref.norm().clamp_min(1e-30), (got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)]
        self.step_e2e(False)  # the same step through the public API (rl_micro_step with main_grad=<the accumulator>)
        torch.cuda.synchronize()
        vals.append((self.acc.shard - ref).norm() / ref.norm().clamp_min(1e-30))
        t = torch.stack(vals).double()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return {"rel_l2": float(t[0]), "max_abs_over_max": float(t[1]), "rel_l2_public_api": float(t[2]),
                "ok": bool(float(t[0]) < 1e-5 and float(t[1]) < 1e-5 and float(t[2]) < 1e-5),
                "what": "every rank's shard after one step of the fused accumulation (raw C-ABI loop, and rl_micro_step) vs local fp32 accumulator + ncclReduceScatter on the same micro-batches (max over ranks)"}

    def step_resident_classic(self):
        """one step with the local fp32 accumulator and the NCCL reduce-scatter (reference for dw_sync_check)"""
        acc, self.acc = self.acc, None
        try:
            self.step_resident(False)
        finally:
            self.acc = acc

    def rs_check(self):
        """multi-GPU correctness in the record: the reduce-scattered d_W_lm shard of every rank equals the sum over ranks of
        that shard of their local accumulators (fp64 checksums, all-reduced)"""
        if self.dist is None or self.rs_out is None:
            return None
        V, H = self.wl["V"], self.wl["H"]
        per = V // self.world
        torch.cuda.synchronize()
        local = self.main_grad.view(self.world, per * H).double()
        sums = torch.stack([local.sum(dim=1), local.abs().sum(dim=1)])  # [2, world]
        self.dist.all_reduce(sums)
        rank = self.dist.get_rank()
        got = self.rs_out.double().sum()
        rel = float((got - sums[0, rank]).abs() / sums[1, rank].clamp_min(1e-300))
        t = torch.tensor([rel], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return {"max_rel_checksum_error": float(t[0]), "ok": bool(float(t[0]) < 1e-6),
                "what": "sum(reduce-scattered shard) vs sum over ranks of sum(local fp32 accumulator shard), relative to the sum of magnitudes"}

    # ---- per-kernel-family timing (one family per launch sequence, CUDA events on the launching stream) ----
    def breakdown(self, save: bool):
        P, wl, dev = self.P, self.wl, self.dev
        H, V, T = wl["H"], wl["V"], wl["T"]
        mb = self.dev_mbs[0]
        hid, ids = mb["hidden"][0], mb["input_ids"][0]
        logp, ent, lse, _ = P.lm_head_forward_raw(hid, self.Wd, ids, 1.0, workspace=self.ws_f)
        dlogp = torch.randn(T, device=dev) / T

        def ev_time(fn, reps):
            """average launch-sequence time over ~1 s of back-to-back repetitions (so the SM clock has settled under the
            power cap, like inside the timed step), CUDA events on the launching stream"""
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n, t0 = 0, time.perf_counter()
            a.record()
            while n < reps or time.perf_counter() - t0 < 1.0:
                fn()
                n += 1
                if n % 8 == 0:
                    torch.cuda.synchronize()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        fam = {}
        if save:
            self.ensure_saved_block()
            vc, nchunks = V, 1
            fwd_s = lambda: P.lm_head_forward_save_raw(hid, self.Wd, ids, 1.0, workspace=self.ws_f, saved=self.saved_blk)  # noqa: E731
            bwd_s = lambda m: P.lm_head_backward_saved_raw(hid, self.Wd, ids, lse, dlogp, 1.0, self.saved_blk, dweight=self.main_grad,  # noqa: E731
                                                           dweight_accumulate=True, phase_mask=m, workspace=self.ws_bs)
            fam["lmhead_fwd_save_kernel"] = dict(ms_per_launch=ev_time(fwd_s, 8), launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            bwd_s(1)  # the block now holds real dlogits (tensor-core power is data dependent), as inside the step
            for name, mask in (("lmhead_bwd_dhidden_kernel", 2), ("lmhead_bwd_dweight_kernel", 4)):
                fam[name] = dict(ms_per_launch=ev_time(lambda: bwd_s(mask), 4), launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            # HBM-bound in-place pass (re-applied to its own output here: timing only)
            fam["dz_from_saved_logits_kernel"] = dict(ms_per_launch=ev_time(lambda: bwd_s(1), 8), launches_per_mb=1, flops_per_launch=0.0,
                                                      bytes_per_launch=algorithmic_bytes("dz_from_saved_logits_kernel", T, H, V, V))
        else:
            vc = min(VCHUNK, (V + 255) // 256 * 256)
            nchunks = (V + vc - 1) // vc
            fam["lmhead_fwd_kernel"] = dict(ms_per_launch=ev_time(lambda: P.lm_head_forward_raw(hid, self.Wd, ids, 1.0, workspace=self.ws_f), 8),
                                            launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            for name, mask in (("lmhead_bwd_dz_kernel", 1), ("lmhead_bwd_dhidden_kernel", 2), ("lmhead_bwd_dweight_kernel", 4)):
                t_ms = ev_time(lambda: P.lm_head_backward_raw(hid, self.Wd, ids, lse, dlogp, 1.0, dweight=self.main_grad,
                                                              dweight_accumulate=True, phase_mask=mask, workspace=self.ws_b), 4)
                fam[name] = dict(ms_per_launch=t_ms / nchunks, launches_per_mb=nchunks, flops_per_launch=2.0 * T * H * V / nchunks)
        return fam, vc


def measure_mode(rn: Runner, save: bool, steps: int, warmup: int, sampler, rank: int, world: int, with_breakdown: bool):
    """resident + end-to-end timing of one backward mode; returns the mode's record"""
    wl, args = rn.wl, rn.args
    H, V, T, n_mb = wl["H"], wl["V"], wl["T"], rn.n_mb
    peaks = measured_peaks()
    mode = "saved" if save else "recompute"
    if sampler is not None:
        sampler.start()
    rn._lib.launch_count(reset=True)
    ms_res, ms_res_local, rs_ms = rn.timed(lambda: rn.step_resident(save), steps, warmup)
    launches = rn._lib.launch_count(reset=True)
    ms_e2e, ms_e2e_local, _ = rn.timed(lambda: rn.step_e2e(save, args.forward_skips_masked), steps, warmup)
    clocks = sampler.stop() if sampler is not None else {}
    tokens_per_step = n_mb * T
    value = world * tokens_per_step * steps / (ms_res * 1e-3)
    e2e_value = world * tokens_per_step * steps / (ms_e2e * 1e-3)
    flops_per_token = 6.0 * H * V
    alg = value / world * flops_per_token / 1e12
    rec = {
        "value": value, "per_gpu": value / world, "ms_per_step": ms_res / steps, "steps": steps, "warmup": warmup,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rn.h2d_bytes, "d2h_bytes_per_step": 4 * n_mb + 8 * n_mb,
                "ms_per_step": ms_e2e / steps, "api": "prime_rl_cse_b200.rl_micro_step + compute_advantages (host buffers)"},
        "gpu_launches": launches * steps // (steps + warmup),
        "clocks": clocks,
        "lm_head_backward": (f"saved-logits (opt-in): one bf16 [T, V] block ({T * V * 2 / 2**30:.2f} GiB) kept per micro-batch between forward and "
                             "backward, rewritten in place into dlogits; 3 GEMMs of T*H*V per micro-batch" if save else
                             f"recompute (default; north_star): lse[T] kept only, logits recomputed per {VCHUNK}-column vocab chunk, the "
                             "[T, V] logits never exist in HBM; 4 GEMMs of T*H*V per micro-batch, 3 credited"),
        "tensor_tflops_algorithmic_per_gpu": alg,
        "frac_of_peak": {"measured_burst": alg / peaks["burst"], "measured_sustained": alg / peaks["sustained"],
                         "nominal_2250": alg / 2250.0, "executed_over_algorithmic": (6.0 if save else 8.0) / 6.0},
    }
    # per-rank table (multi-GPU): this rank's own step time, its d_W reduce-scatter time, its SM clock under load
    if rn.dist is not None:
        mine = torch.tensor([ms_res_local / steps, rs_ms if rs_ms is not None else -1.0, clocks.get("sm_mhz") or -1.0, ms_e2e_local / steps],
                            dtype=torch.float64, device=rn.dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        rn.dist.all_gather(allr, mine)
        tab = torch.stack(allr).cpu()
        rec["per_rank"] = {"ms_per_step": [round(float(x), 3) for x in tab[:, 0]], "dw_reduce_scatter_ms": [round(float(x), 3) for x in tab[:, 1]],
                           "sm_mhz": [float(x) for x in tab[:, 2]], "e2e_ms_per_step": [round(float(x), 3) for x in tab[:, 3]],
                           "note": "ms_per_step: each rank's own CUDA-event time over the timed region / steps (includes waiting for the "
                                   "slowest rank in the step-end d_W_lm sync: the NCCL reduce-scatter, or the barrier that is left of it when "
                                   "the gradient is accumulated in the reduce-scattered layout); value uses the max"}
        fused = rn.acc is not None and not save
        rec["dw_sync"] = rn.acc_note if fused else "local fp32 accumulator + ncclReduceScatter once per step"
        if fused:
            rec["dw_sync_check"] = rn.dw_sync_check()
            rn.main_grad.zero_()
        else:
            rec["dw_reduce_scatter_check"] = rn.rs_check()
    if with_breakdown and rank == 0:
        fam, vc = rn.breakdown(save)
        top = max(fam, key=lambda k: fam[k]["ms_per_launch"] * fam[k]["launches_per_mb"])
        f = fam[top]
        achieved = f["flops_per_launch"] / (f["ms_per_launch"] * 1e-3) / 1e12
        tr = ncu_traffic(mode, top, wl["shape"])
        rec["roofline"] = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                           "frac": achieved / peaks["sustained"], "traffic": tr["bytes"] if tr else None,
                           "traffic_source": tr["source"] if tr else "no ncu capture committed for this workload/kernel",
                           "algorithmic_bytes": algorithmic_bytes(top, T, H, V, vc),
                           "flops_per_launch": f["flops_per_launch"], "ms_per_launch": f["ms_per_launch"],
                           "peak_source": peaks["source"] + " sustained (kernel timed in a ~1 s back-to-back loop under the power cap)",
                           "frac_of_burst": achieved / peaks["burst"]}
        rec["kernel_breakdown_ms_per_micro_batch"] = {k: v["ms_per_launch"] * v["launches_per_mb"] for k, v in fam.items()}
        rec["kernel_tflops"] = {k: v["flops_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e12 for k, v in fam.items() if v["flops_per_launch"]}
        hbm = {k: v for k, v in fam.items() if v.get("bytes_per_launch")}
        if hbm:  # the HBM-bound pass of the saved-logits backward against the measured copy bandwidth
            rec["hbm_kernels"] = {k: {"achieved_gbs": v["bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9, "peak_gbs": peaks["hbm"],
                                      "frac": v["bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9 / peaks["hbm"],
                                      "algorithmic_bytes": v["bytes_per_launch"]} for k, v in hbm.items()}
    return rec


REGIMES = [
    # name, shape, micro-batches per step, regime overrides
    ("d_peaked", "qwen3-8b-hendrycks", 4, dict(dist="peaked")),
    ("tau_0.6", "qwen3-8b-hendrycks", 4, dict(tau=0.6)),
    ("sigma_1.5_clipping", "qwen3-8b-hendrycks", 4, dict(dist="peaked", sigma=1.5)),
    ("masked_prompt_25pct", "qwen3-8b-hendrycks", 4, dict(prompt_frac=0.25)),
    ("multi_turn_wordle_1.7b", "qwen3-1.7b-wordle", 4, dict(prompt_frac=0.25, multi_turn=True)),
    ("qwen3_1.7b_wordle_all_ones", "qwen3-1.7b-wordle", 4, {}),
    ("sequence_mode_deepscaler_32k", "qwen3-8b-deepscaler", 2, dict(tau=0.6)),
    ("qwen3_0.6b_reverse_config1_shape", "qwen3-0.6b-reverse", 4, {}),
]


def run_regimes(args, dev, names=None):
    """short end-to-end runs (public API, host buffers, recompute backward) of the other regimes of SURVEY §8(d): 1 warm-up +
    2 timed steps each; not the headline, but in the driver's record"""
    out = {}
    for name, shape, n_mb, rg in REGIMES:
        if names and name not in names:
            continue
        try:
            wl = make_workload(shape, 0, n_distinct=2, pin=True, device=dev, regime=rg, n_mb=n_mb)
            rn = Runner(wl, dev, None, 1, args)
            ms, _, _ = rn.timed(lambda: rn.step_e2e(False), 2, 1)
            mb = rn.dev_mbs[0]
            rec = {"e2e_value": wl["n_mb"] * wl["T"] * 2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 2, "workload": describe(wl),
                   "lm_head_backward": "recompute", "masked_fraction": 1.0 - float(mb["loss_mask"].float().mean())}
            if wl["ratio_type"] == "token":  # how much clipping the regime exercises
                out_d = rn.P.rl_micro_step(mb["hidden"], rn.W, mb, rn.cfg, rn.loss_scale, backward=False)
                rec["clipped_fraction"] = float(out_d["is_clipped"].float().mean())
                rec["mean_entropy_nats"] = float(out_d["entropy"].mean())
            rec["frac_of_measured_burst"] = rec["e2e_value"] * 6.0 * wl["H"] * wl["V"] / 1e12 / measured_peaks()["burst"]
            out[name] = rec
            del rn, wl
        except Exception as e:  # noqa: BLE001  context records only: never fatal for the headline
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import prime_rl_cse_b200 as P  # noqa: F401
    from prime_rl_cse_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    # the vocab-sharded workload also runs as a single torchrun rank (one shard = the whole vocabulary): its exchange steps,
    # the peer-memory epilogue and the ragged-row padding are then exercised on a one-GPU box too
    if world > 1 or (args.workload == "qwen3-32b-vocab8" and "RANK" in os.environ):
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)
    _lib.load()
    if args.workload == "qwen3-32b-vocab8":
        res = run_vocab_sharded(args, rank, world, dev, dist)
        if rank == 0:
            res.update({"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                        "vs_baseline": None, "dtype": "bf16", "data": "synthetic"})
            emit(res)
        dist.barrier()
        dist.destroy_process_group()
        return

    regime = dict(prompt_frac=args.prompt_frac)
    wl = make_workload(args.workload, rank, n_distinct=args.distinct, pin=True, device=dev, regime=regime, n_mb=args.micro_batches)
    rn = Runner(wl, dev, dist, world, args)
    H, V, T, n_mb = wl["H"], wl["V"], wl["T"], rn.n_mb
    sampler = ClockSampler(local)  # every rank samples its own GPU
    modes = ["recompute", "saved"] if args.lm_head_backward == "both" else [args.lm_head_backward]
    recs = {}
    for m in modes:
        # the headline mode is timed over EXACTLY --steps steps after --warmup; the labelled sub-record of the other mode over at
        # most 5 steps after at most 3 warm-ups (stated in the record), so that a long --steps run does not double its wall time
        k, wu = (args.steps, args.warmup) if m == modes[0] else (min(args.steps, 5), min(args.warmup, 3))
        recs[m] = measure_mode(rn, m == "saved", k, wu, sampler, rank, world, not args.no_breakdown)
        if m == "saved":
            rn.saved_blk = None
            torch.cuda.empty_cache()
    head_mode = modes[0]
    head = recs[head_mode]
    result = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": describe(wl, world), "tokens_per_gpu_per_step": n_mb * T,
                   "parallelism": f"token-sharded dp{world} (no data-path collective; "
                   + (f"d_W_lm gradient sync: {head.get('dw_sync')})" if world > 1 else "single GPU)"),
                   "l2_policy": "inputs larger than L2 (W_lm 1.24 GB + 134 MB hidden per micro-batch, rotating micro-batches)",
                   "flops_per_token_algorithmic": 6.0 * H * V, "lm_head_backward": head["lm_head_backward"],
                   "headline_mode": head_mode},
        "per_gpu": head["per_gpu"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
        "tensor_tflops_algorithmic_per_gpu": head["tensor_tflops_algorithmic_per_gpu"], "frac_of_peak": head["frac_of_peak"],
    }
    for k in ("roofline", "kernel_breakdown_ms_per_micro_batch", "kernel_tflops", "hbm_kernels", "per_rank", "dw_reduce_scatter_check", "dw_sync",
              "dw_sync_check"):
        if k in head:
            result[k] = head[k]
    result["modes"] = {("saved_logits" if m == "saved" else m): recs[m] for m in modes}
    if rn.acc is not None:
        rn.acc.close()
        rn.acc = None
    if dist is not None and not args.no_vocab_check:
        import prime_rl_cse_b200 as P_

        try:
            result["vocab_sharded_check"] = vocab_sharded_check(P_, dist, rn, world, rank, dev)
        except Exception as e:  # noqa: BLE001  a context record: never fatal for the headline (every rank takes the same path)
            result["vocab_sharded_check"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    del rn
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_regimes and args.workload == "qwen3-8b-hendrycks":
        result["regimes"] = run_regimes(args, dev)
    if rank == 0 and not args.no_torch_baseline:
        result["torch_gpu_reference"] = torch_gpu_reference(wl, dev)
    if rank == 0:
        if not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(args.workload, steps=3, warmup=0)
            if args.workload == "qwen3-8b-hendrycks" and world == 1:
                result["cpu_baseline_config1"] = cpu_baseline_config1()
        emit(result)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: Qwen3-32B shape, LM head vocab-sharded over the ranks (torchrun x N)
# ------------------------------------------------------------------------------------------------
def run_vocab_sharded(args, rank, world, dev, dist):
    """Every rank owns V/world rows of W_lm (FSDP2's dim-0 shard, model.py:100-135) and brings its own packed micro-batch per
    round; see prime_rl_cse_b200.parallel.VocabShardedLMHead for the exchange steps.  Weak scaling in N: tokens per rank per
    step are fixed, the shard shrinks as 1/N, so the work per rank per round (N*T tokens x V/N columns) is constant."""
    import prime_rl_cse_b200 as P
    from prime_rl_cse_b200.parallel import VocabShardedLMHead

    if dist is None:
        raise SystemExit("--workload qwen3-32b-vocab8 runs under torchrun (python bench.py --gpus N --workload qwen3-32b-vocab8, N >= 2; "
                         "or one rank: python -m torch.distributed.run --nproc-per-node 1 bench.py --workload qwen3-32b-vocab8)")
    H, V, T, nseg, n_mb, group, ratio_type = SHAPES[args.workload]
    n_mb = args.micro_batches if args.micro_batches > 0 else n_mb
    peaks = measured_peaks()
    g = torch.Generator(device=dev).manual_seed(4321)  # the SAME W on every rank (each keeps its shard; the full copy is for the check)
    W_full = (torch.randn(V, H, generator=g, dtype=torch.float32, device=dev) * 0.02).to(torch.bfloat16)
    v0, v1 = P.vocab_shard_bounds(V, world, rank)
    W_shard = W_full[v0:v1].clone()
    main_grad = torch.zeros(v1 - v0, H, dtype=torch.float32, device=dev)
    head = VocabShardedLMHead(W_shard, V, main_grad=main_grad, fused_rs=bool(args.fused_rs))
    cfg = P.LossConfig(ratio_type=ratio_type, ratio_length_norm=False, clip_ratio=8.0)
    gl = torch.Generator(device=dev).manual_seed(1234 + rank)  # different tokens on every rank
    seg = [T // nseg] * nseg
    mbs = []
    for i in range(min(args.distinct, n_mb)):
        hidden = torch.randn(T, H, generator=gl, device=dev).to(torch.bfloat16)
        ids = torch.randint(0, V, (T,), generator=gl, device=dev)
        lp, _, _, _ = P.lm_head_forward_raw(hidden, W_full, ids, 1.0)
        old = lp + 0.15 * torch.randn(T, generator=gl, device=dev)
        adv = torch.cat([torch.full((n,), float(a)) for n, a in zip(seg, (0.5, -0.25, 1.0, -0.5))]).to(dev)
        mb = dict(hidden=hidden.unsqueeze(0), input_ids=ids.unsqueeze(0), position_ids=torch.cat([torch.arange(n) for n in seg]).unsqueeze(0).to(dev),
                  advantages=adv.unsqueeze(0), logprobs=old.unsqueeze(0), loss_mask=torch.ones(1, T, dtype=torch.bool, device=dev), temperature=1.0)
        mbs.append(mb)
    host_mbs = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in mb.items()} for mb in mbs]
    loss_scale = n_mb * T
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.empty(n_mb, dtype=torch.float32).pin_memory()

    parity = vocab_sharded_parity(P, dist, head, mbs[0], W_full, seg, world, dev)
    parity["ragged_rows"] = vocab_sharded_parity(P, dist, head, mbs[0], W_full, seg, world, dev, ragged=True)
    parity["ok"] = bool(parity["ok"] and parity["ragged_rows"]["ok"])
    main_grad.zero_()
    torch.cuda.empty_cache()

    def step(e2e: bool):
        cur = torch.cuda.current_stream()

        def stage(i):
            if not e2e:
                mb = mbs[i % len(mbs)]
            else:
                with torch.cuda.stream(copy_stream):
                    mb = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host_mbs[i % len(host_mbs)].items()}
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                cur.wait_event(ev)
                for v in mb.values():
                    if torch.is_tensor(v):
                        v.record_stream(cur)
            return mb, head.gather_inputs(mb["hidden"], mb["input_ids"])  # the all-gather runs ahead on the comm stream

        nxt = stage(0)
        for i in range(n_mb):
            mb, gathered = nxt
            if i + 1 < n_mb:
                nxt = stage(i + 1)
            o = head.micro_step(mb["hidden"], mb, cfg, loss_scale, gathered=gathered)
            if e2e:
                loss_host[i].copy_(o["loss"], non_blocking=True)
        if e2e:
            torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        mine = float(t[0])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), mine

    sampler = ClockSampler(dev.index)
    sampler.start()
    from prime_rl_cse_b200 import _lib

    _lib.launch_count(reset=True)
    ms_res, ms_mine = timed(lambda: step(False), args.steps, args.warmup)
    launches = _lib.launch_count(reset=True)
    exch = head.exchange_ms()
    ms_e2e, _ = timed(lambda: step(True), args.steps, args.warmup)
    clocks = sampler.stop()
    tab = torch.tensor([ms_mine / args.steps, clocks.get("sm_mhz") or -1.0], dtype=torch.float64, device=dev)
    allr = [torch.empty_like(tab) for _ in range(world)]
    dist.all_gather(allr, tab)
    tokens = world * n_mb * T
    value = tokens * args.steps / (ms_res * 1e-3)
    alg = value / world * 6.0 * H * V / 1e12
    return {
        "value": value, "ms_per_step": ms_res / args.steps, "scaling": "weak", "per_gpu": value / world,
        "config": {"workload": f"{args.workload}: H={H} V={V} vocab-sharded {world}-way ({v1 - v0} rows per rank), {n_mb} rounds per step, every rank "
                               f"brings one packed micro-batch of [1,{T}] tokens per round ({world * T} tokens swept per rank per round), "
                               "token-ratio loss, all-ones mask, tau=1, D-uniform",
                   "parallelism": f"vocab-sharded tp{world}: all-gather hidden, all-to-all of 16 B/token partials, all-gather (lse, dlogp), "
                                  + ("d_hidden reduce-scatter fused into the GEMM epilogue (TMA reduce-add into the owner's symmetric memory)"
                                     if head.fused_rs else "NCCL reduce-scatter of d_hidden after the GEMM"),
                   "lm_head_backward": "recompute", "flops_per_token_algorithmic": 6.0 * H * V,
                   "l2_policy": "inputs larger than L2 (hidden of a round: 1.3 GB at 8 ranks)"},
        "e2e": {"value": tokens * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_mb * sum(v.numel() * v.element_size() for v in host_mbs[0].values() if torch.is_tensor(v)),
                "d2h_bytes_per_step": 4 * n_mb, "ms_per_step": ms_e2e / args.steps, "api": "VocabShardedLMHead.micro_step (host buffers)"},
        "gpu_launches": launches * args.steps // (args.steps + args.warmup), "clocks": clocks,
        "exchange_ms_last_round": exch,
        "parity_vs_unsharded": parity,
        "per_rank": {"ms_per_step": [round(float(x[0]), 3) for x in allr], "sm_mhz": [float(x[1]) for x in allr]},
        "tensor_tflops_algorithmic_per_gpu": alg,
        "frac_of_peak": {"measured_burst": alg / peaks["burst"], "measured_sustained": alg / peaks["sustained"], "nominal_2250": alg / 2250.0,
                         "executed_over_algorithmic": 8.0 / 6.0},
        "roofline": {"kernel": "the four GEMM sweeps of a round (forward, dlogits recompute, d_hidden, d_weight) on one rank", "bound": "tensor",
                     "achieved": alg * 8.0 / 6.0, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": alg * 8.0 / 6.0 / peaks["sustained"],
                     "traffic": None, "note": "executed flops (8 H V per token) / step time, exchanges included"},
    }


def vocab_sharded_parity(P, dist, head, mb, W_full, seg, world, dev, ragged: bool = False):
    """One round of the vocab-sharded head from a zeroed d_W accumulator against the UNSHARDED fused path on this GPU (same
    kernels, full W): every rank checks its own tokens' log-probs / entropy / d_hidden and its d_W shard; max over ranks.
    `ragged`: every rank keeps a different number of tokens of its row (packed rows are not padded, batch.py:115-116) and the
    head pads them to the common length (`pad_to`)."""
    from prime_rl_cse_b200.loss import _loss_packed_raw

    T, V, H = mb["input_ids"].shape[1], W_full.shape[0], W_full.shape[1]
    n, pad_to = T, None
    if ragged:
        n, pad_to = max(T - 96 * (dist.get_rank() + 1) - 5, 2), T
        mb = {k: (v[:, :n].contiguous() if torch.is_tensor(v) else v) for k, v in mb.items()}
        mb["position_ids"] = torch.arange(n, device=dev).unsqueeze(0)
        seg = [n]
    tau = float(mb["temperature"])
    head.main_grad.zero_()
    cfg = P.LossConfig(ratio_type="token", ratio_length_norm=False, clip_ratio=8.0)
    out = head.micro_step(mb["hidden"], mb, cfg, n, pad_to=pad_to)
    gathered = head.gather_inputs(mb["hidden"], mb["input_ids"], pad_to)
    torch.cuda.current_stream().wait_event(gathered["ready"])
    lp_u, ent_u, lse_u, _ = P.lm_head_forward_raw(mb["hidden"][0], W_full, mb["input_ids"][0], 1.0 / tau)
    cu = torch.tensor([0] + list(torch.tensor(seg).cumsum(0)), dtype=torch.int32, device=dev)
    _, _, _, _, g_u = _loss_packed_raw(lp_u, mb["logprobs"][0], mb["advantages"][0], mb["loss_mask"][0].view(torch.uint8), cu, len(seg), 0,
                                       False, 8.0, n)
    dh_u, _ = P.lm_head_backward_raw(mb["hidden"][0], W_full, mb["input_ids"][0], lse_u, g_u, 1.0 / tau, want_dweight=False)
    # unsharded d_W needs every rank's tokens: gather (lse, dlogp) of all ranks, run the unsharded backward over world*T tokens
    # (padded positions of a ragged row: zero hidden rows with dlogp = 0, which add exact zeros)
    pair = torch.zeros(T, 2, device=dev)
    pair[:, 0] = math.log(V)
    pair[:n, 0] = lse_u
    pair[:n, 1] = g_u
    pair[0, 1] = 0.0
    pair_all = torch.empty(world * T, 2, device=dev)
    dist.all_gather_into_tensor(pair_all, pair)
    dw_u = torch.zeros(V, H, dtype=torch.float32, device=dev)
    P.lm_head_backward_raw(gathered["h_all"], W_full, gathered["i_all"], pair_all[:, 0].contiguous(), pair_all[:, 1].contiguous(), 1.0 / tau,
                           want_dhidden=False, dweight=dw_u, dweight_accumulate=True)

    def rel(a, b):
        return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))

    par = torch.tensor([float((out["logprobs"] - lp_u).abs().max()), float((out["entropy"] - ent_u).abs().max()), rel(out["d_hidden"], dh_u),
                        rel(head.main_grad, dw_u[head.v0:head.v1])], dtype=torch.float64, device=dev)
    dist.all_reduce(par, op=dist.ReduceOp.MAX)
    return {"max_abs_dlogp": float(par[0]), "max_abs_dentropy": float(par[1]), "rel_l2_d_hidden": float(par[2]), "rel_l2_d_weight_shard": float(par[3]),
            "ok": bool(par[0] < 1e-3 and par[1] < 1e-3 and par[2] < 1e-2 and par[3] < 1e-2),
            "what": f"{world}-rank vocab-sharded round vs the unsharded fused path on each rank's own tokens (max over ranks)"
                    + (f"; ragged rows: rank r keeps {T} - 96 (r + 1) - 5 tokens, padded to {T} inside the head" if ragged else "")}


def vocab_sharded_check(P, dist, rn, world, rank, dev, T=2048):
    """Every N > 1 line of the default (token-sharded) workload also carries a SMALL vocab-sharded round at the same H and V
    (T tokens per rank), so that config 4's exchange steps and the fused epilogue reduce-scatter are in the driver's own
    multi-GPU record: parity against the unsharded path and the round time, NCCL and fused."""
    from prime_rl_cse_b200.parallel import VocabShardedLMHead

    W_full = rn.Wd.clone()
    dist.broadcast(W_full, src=0)  # the token-sharded workload seeds W per rank; the vocab-sharded head needs ONE W
    H, V = W_full.shape[1], W_full.shape[0]
    v0, v1 = P.vocab_shard_bounds(V, world, rank)
    g = torch.Generator(device=dev).manual_seed(777 + rank)
    seg = [T // 4] * 4
    hidden = torch.randn(T, H, generator=g, device=dev).to(torch.bfloat16)
    ids = torch.randint(0, V, (T,), generator=g, device=dev)
    lp, _, _, _ = P.lm_head_forward_raw(hidden, W_full, ids, 1.0)
    mb = dict(hidden=hidden.unsqueeze(0), input_ids=ids.unsqueeze(0), position_ids=torch.cat([torch.arange(n) for n in seg]).unsqueeze(0).to(dev),
              advantages=torch.cat([torch.full((n,), a) for n, a in zip(seg, (0.5, -0.25, 1.0, -0.5))]).unsqueeze(0).to(dev),
              logprobs=(lp + 0.15 * torch.randn(T, generator=g, device=dev)).unsqueeze(0), loss_mask=torch.ones(1, T, dtype=torch.bool, device=dev),
              temperature=1.0)
    out = {}
    for fused in (False, True):
        head = VocabShardedLMHead(W_full[v0:v1].clone(), V, main_grad=torch.zeros(v1 - v0, H, dtype=torch.float32, device=dev), fused_rs=fused)
        rec = vocab_sharded_parity(P, dist, head, mb, W_full, seg, world, dev)
        rag = vocab_sharded_parity(P, dist, head, mb, W_full, seg, world, dev, ragged=True)  # unpadded packed rows of different lengths
        rec["ragged_rows"] = {k: rag[k] for k in ("max_abs_dlogp", "max_abs_dentropy", "rel_l2_d_hidden", "rel_l2_d_weight_shard", "ok")}
        rec["ok"] = bool(rec["ok"] and rag["ok"])
        head.main_grad.zero_()
        cfg = P.LossConfig()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            head.micro_step(mb["hidden"], mb, cfg, T)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec["ms_per_round"] = float(t[0])
        rec["tokens_per_rank_per_round"] = T
        out["fused_epilogue_reduce_scatter" if fused else "nccl_reduce_scatter"] = rec
        head.close()
        del head
    return out


def torch_gpu_reference(wl, dev, T_s=4096, reps=3):
    """The reference's own GPU path for the same micro-step (bf16 lm_head GEMM via cuBLAS -> fp32 logits -> shift -> /tau ->
    log-softmax gather -> loss -> entropy -> autograd backward; oracle/torch_port.py restating rl/train.py:285-310) on this
    B200, on a 4096-token slice of the workload (materialised [T,V] fp32 logits: ~2.5 GB per temporary).  Context only."""
    from oracle import torch_port

    mb = wl["mbs"][0]
    T_s = min(T_s, wl["T"])
    nseg = 4
    seg = [T_s // nseg] * nseg
    hidden0 = mb["hidden"][0, :T_s].to(dev)
    weight0 = wl["weight"].to(dev)
    ids = mb["input_ids"][0, :T_s].to(dev)
    pos = torch.cat([torch.arange(n) for n in seg]).to(dev)
    adv = mb["advantages"][0, :T_s].to(dev)
    mask = torch.ones(T_s, dtype=torch.bool, device=dev)
    old = mb["logprobs"][0, :T_s].to(dev)

    def one():
        h = hidden0.clone().requires_grad_(True)
        w = weight0.detach().clone().requires_grad_(True)
        torch_port.micro_step_cpu(h, w, ids, pos, old, adv, mask, 1.0, wl["ratio_type"], False, 8.0, T_s)

    try:
        one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            one()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        peak = torch.cuda.max_memory_allocated(dev) / 2**30
        return {"value": T_s / ms * 1e3, "unit": UNIT, "sample": f"{T_s}-token packed row of the same shape, torch ops on this GPU "
                f"(cuBLAS bf16 lm_head + ATen, materialised fp32 logits), mean of {reps}", "ms_per_sample": ms, "peak_mem_gib": peak}
    except Exception as e:  # noqa: BLE001  (e.g. out of memory on a shared box): context only, never fatal
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    finally:
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own torch-CPU path (oracle/torch_port.py) on host cores
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_TOKENS = 2048


def cpu_inputs(wl, T_s, nseg=4):
    """leaf tensors and per-token inputs of one bounded CPU sample, prepared OUTSIDE any timed region"""
    from oracle import torch_port

    mb = wl["mbs"][0]
    seg = [T_s // nseg] * nseg
    seg[-1] += T_s - sum(seg)
    adv_list = torch_port.advantages_cpu(wl["rewards"].tolist(), wl["lengths"].tolist(), wl["group"], "drgrpo")
    return dict(hidden=mb["hidden"][0, :T_s].float().clone(),  # config 1 runs fp32 on the CPU
                weight=wl["weight"].float().clone(), ids=mb["input_ids"][0, :T_s], pos=torch.cat([torch.arange(n) for n in seg]),
                adv=torch.cat([torch.full((n,), adv_list[i % len(adv_list)]) for i, n in enumerate(seg)]),
                mask=torch.ones(T_s, dtype=torch.bool), old=mb["logprobs"][0, :T_s].clone(), T=T_s)


def cpu_step(wl, inp):
    """one timed CPU micro-step: train.py:285-310 (forward, loss, entropy, backward) + the step's advantages.  Returns seconds."""
    from oracle import torch_port

    hidden = inp["hidden"].detach().requires_grad_(True)
    weight = inp["weight"].detach().requires_grad_(True)
    hidden.grad = weight.grad = None
    t0 = time.perf_counter()
    torch_port.advantages_cpu(wl["rewards"].tolist(), wl["lengths"].tolist(), wl["group"], "drgrpo")
    torch_port.micro_step_cpu(hidden, weight, inp["ids"], inp["pos"], inp["old"], inp["adv"], inp["mask"], 1.0, wl["ratio_type"],
                              wl["ratio_type"] == "sequence", 8.0, inp["T"])
    return time.perf_counter() - t0


def cpu_baseline(workload: str, steps: int, warmup: int) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(workload, 0, n_distinct=1, pin=False)
    T_s = min(CPU_SAMPLE_TOKENS, wl["T"])
    inp = cpu_inputs(wl, T_s)
    for _ in range(warmup):
        cpu_step(wl, inp)
    times = [cpu_step(wl, inp) for _ in range(max(steps, 1))]
    dt = statistics.median(times)
    return {"value": T_s / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one packed row of {T_s} tokens (4 segments) at H={wl['H']} V={wl['V']}, fp32, torch CPU ops "
                      f"(oracle/torch_port.py restating rl/train.py:285-310), median of {max(steps,1)} run(s), {dt:.2f} s each"}


def cpu_baseline_config1() -> dict:
    """BASELINE config 1 EXACTLY: Qwen3-0.6B shape, 4 problems x 8 rollouts x 128 tokens = one packed
    [1, 4096] row, fp32, all host cores, compute_advantages + RL loss fwd/bwd; 1 warm-up + 3 timed runs, median."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload("qwen3-0.6b-reverse", 0, n_distinct=1, pin=False)
    inp = cpu_inputs(wl, wl["T"], nseg=32)
    cpu_step(wl, inp)
    times = [cpu_step(wl, inp) for _ in range(3)]
    dt = statistics.median(times)
    return {"value": wl["T"] / dt, "unit": UNIT, "cores": torch.get_num_threads(), "os_cpu_count": cores, "kind": "port",
            "sample": f"config 1 exactly: T={wl['T']} (32 x 128-token rollouts, 4 groups of 8), H={wl['H']}, V={wl['V']}, fp32, "
                      f"1 warm-up + 3 timed, median {dt:.2f} s per step"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU op sequence for the path (oracle/torch_port.py: the reference is pure Python
    and /root/reference does not travel to the GPU box) on the host cores, each step a bounded sample of the B200 arm's
    workload; inputs are prepared once, outside the timed loop."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(args.workload, 0, n_distinct=1, pin=False)
    T_s = min(CPU_SAMPLE_TOKENS, wl["T"])
    inp = cpu_inputs(wl, T_s)
    for _ in range(args.warmup):
        cpu_step(wl, inp)
    dt = sum(cpu_step(wl, inp) for _ in range(args.steps))
    value = T_s * args.steps / dt
    H, V = wl["H"], wl["V"]
    sample = (f"each step = one packed row of {T_s} tokens (4 segments) of the same workload shape, fp32, all host threads "
              f"(torch CPU ops, oracle/torch_port.py restating the reference's rl/train.py:285-310; the reference is pure Python "
              f"and /root/reference does not travel to the GPU box); leaf-tensor setup is outside the timed region")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: H={H} V={V}, bounded sample of {T_s} tokens per step (the B200 arm runs "
                               f"{wl['n_mb'] * wl['T']} tokens per step of the same shape)", "parallelism": "host CPU, rank 0 only",
                   "same_config_as_b200_arm": False},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qwen3-8b-hendrycks", choices=sorted(SHAPES))
    ap.add_argument("--micro-batches", type=int, default=0, help="override micro-batches per step (0 = the workload's)")
    ap.add_argument("--distinct", type=int, default=3, help="distinct micro-batches kept resident and rotated")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the short runs of the other measurement regimes (N=1 only)")
    ap.add_argument("--no-vocab-check", action="store_true", help="N>1: skip the small vocab-sharded round (config 4 parity + round time)")
    ap.add_argument("--forward-skips-masked", action="store_true", help="e2e arm: also restrict the forward to unmasked positions")
    ap.add_argument("--prompt-frac", type=float, default=0.0, help="mask the first fraction of every rollout (default 0 = all-ones mask)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the torch-on-GPU reference-chain timing")
    ap.add_argument("--lm-head-backward", default="both", choices=["both", "saved", "recompute"],
                    help="both (default): time the recompute backward (headline) and the saved-logits backward (modes.saved_logits); "
                         "recompute / saved: only that one")
    ap.add_argument("--fused-dw-sync", type=int, default=1, help="N>1, recompute mode: 1 = the last micro-batch's d_weight GEMM epilogue sends "
                    "(local total + its tiles) into the owners' shards over peer memory (no step-end collective), 2 = every micro-batch adds "
                    "straight into the owners' shards (no [V, H] gradient per rank), 0 = local accumulator + ncclReduceScatter")
    ap.add_argument("--fused-rs", type=int, default=1, help="vocab-sharded workload: 1 = d_hidden reduce-scatter fused into the GEMM "
                    "epilogue over peer memory, 0 = NCCL reduce-scatter after the GEMM")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_b200(args)


if __name__ == "__main__":
    main()
