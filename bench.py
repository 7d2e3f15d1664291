#!/usr/bin/env python
"""bench.py — PRIME-RL per-token RL loss hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one rank's share of a training batch at BASELINE.json's
Qwen3-8B shape (configs[2], the config the metric is quoted on): 32 packed micro-batches of [1, 16384]
tokens (4 x 4096-token rollouts each), hidden 4096, vocab 151936, bf16 -- per micro-batch: fused LM-head
forward (log-probs + entropy), packed clipped-ratio loss, recompute-based backward to d_hidden and
d_W_lm (fp32 accumulation across micro-batches), plus the step's group advantages.  Weak scaling: every
rank processes its own 32 micro-batches (token shards = FSDP ranks); for N > 1 the accumulated d_W_lm is
reduce-scattered once per step over NCCL as FSDP2 does.

Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
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
    "tiny": (128, 2048, 512, 4, 2, 4, "token"),  # CI smoke of this script only, not a benchmark
}


# DRAM bytes (read + write) per launch of each kernel family at the Qwen3-8B shape (T=16384, H=4096), from `ncu --set full`
# captures (dram__bytes_read.sum + dram__bytes_write.sum).  "recompute": vocab chunk 16384, profiles/r1_lmhead_final_ncu_summary.txt
# (the forward capture there covers a 32768-row vocabulary slice; it is scaled to V=151936).  "saved": whole-vocabulary
# launches, profiles/r1_saved_logits_ncu_summary.txt.
NCU_TRAFFIC_BYTES = {
    "recompute": {
        "lmhead_bwd_dz_kernel": 2.81e9 + 0.508e9,
        "lmhead_bwd_dhidden_kernel": 2.05e9 + 0.245e9,
        "lmhead_bwd_dweight_kernel": 2.06e9 + 0.242e9,
        "lmhead_fwd_kernel": (4.75e9 + 0.007e9) * 151936 / 32768,
    },
    "saved": {
        "lmhead_fwd_save_kernel": 7.834e9 + 4.972e9,
        "dz_from_saved_logits_kernel": 4.979e9 + 4.921e9,
        "lmhead_bwd_dhidden_kernel": 23.650e9 + 0.128e9,
        "lmhead_bwd_dweight_kernel": 18.218e9 + 2.488e9,
    },
}
NCU_TRAFFIC_SOURCE = {"recompute": "profiles/r1_lmhead_final_ncu_summary.txt", "saved": "profiles/r1_saved_logits_ncu_summary.txt"}


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
def make_workload(shape: str, rank: int, n_distinct: int, pin: bool, device="cpu", prompt_frac: float = 0.0):
    """Micro-batches of the named shape (SURVEY §8d "D-uniform": W ~ N(0, 0.02^2), hidden ~ N(0,1)), generated on
    `device` (the GPU for the B200 arm: seconds instead of a single-threaded CPU randn under torchrun) and returned as
    HOST tensors (pinned when `pin`), exactly what a data loader hands the trainer; `weight` stays on `device`."""
    H, V, T, nseg, n_mb, group, ratio_type = SHAPES[shape]
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
        hidden = torch.randn(T, H, generator=g, dtype=torch.float32, device=device).to(torch.bfloat16).cpu()
        ids = torch.randint(0, V, (T,), generator=g, device=device).cpu()
        pos = torch.cat([torch.arange(n) for n in seg])
        mask = torch.ones(T, dtype=torch.bool)  # headline: all ones, as FakeDataLoader (rl/data.py:53)
        old = (-11.9 + 0.15 * torch.randn(T, generator=g, device=device)).cpu()  # near ln(1/V): ratios ~ 1 (near-uniform softmax)
        if prompt_frac > 0:  # --prompt-frac: the first part of every rollout is a prompt (mask 0, old log-prob 0, batch.py:41-42)
            o = 0
            for n in seg:
                mask[o : o + int(n * prompt_frac)] = False
                o += n
            old[~mask] = 0.0
        mb = dict(hidden=hidden.unsqueeze(0), input_ids=ids.unsqueeze(0), position_ids=pos.unsqueeze(0),
                  advantages=torch.cat([adv_r[(i * nseg + j) % n_rollouts].repeat(n) for j, n in enumerate(seg)]).unsqueeze(0),
                  logprobs=old.unsqueeze(0), loss_mask=mask.unsqueeze(0), temperature=1.0)
        if pin:
            for k, v in mb.items():
                if torch.is_tensor(v):
                    mb[k] = v.pin_memory()
        mbs.append(mb)
    return dict(H=H, V=V, T=T, seg=seg, n_mb=n_mb, group=group, ratio_type=ratio_type, weight=weight, mbs=mbs,
                rewards=rewards, lengths=lengths)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import prime_rl_cse_b200 as P
    from prime_rl_cse_b200 import _lib
    from prime_rl_cse_b200.loss import _loss_packed_raw

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)
    _lib.load()

    wl = make_workload(args.workload, rank, n_distinct=args.distinct, pin=True, device=dev, prompt_frac=args.prompt_frac)
    H, V, T, n_mb = wl["H"], wl["V"], wl["T"], wl["n_mb"] if args.micro_batches <= 0 else args.micro_batches
    cfg = P.LossConfig(ratio_type=wl["ratio_type"], ratio_length_norm=False, clip_ratio=8.0)
    mode = 0 if cfg.ratio_type == "token" else 1
    W = wl["weight"].to(dev).requires_grad_(True)  # lm_head.weight is a trainable parameter
    main_grad = torch.zeros(V, H, dtype=torch.float32, device=dev)  # fp32 accumulation across micro-batches
    rs_out = torch.empty((V // world) * H, dtype=torch.float32, device=dev) if world > 1 and V % world == 0 else None
    dev_mbs = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in mb.items()} for mb in wl["mbs"]]
    rewards_d, lengths_d = wl["rewards"].to(dev), wl["lengths"].to(dev)
    bounds = [0]
    for n in wl["seg"]:
        bounds.append(bounds[-1] + n)
    cu = torch.tensor(bounds, dtype=torch.int32, device=dev)
    nseg = len(wl["seg"])
    n_unmasked = int(wl["mbs"][0]["loss_mask"].sum())
    loss_scale = n_mb * n_unmasked if mode == 0 else n_mb  # rl/train.py:263-267
    lib = _lib.load()
    ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
    ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 0), dtype=torch.uint8, device=dev)
    # backward mode: "saved" keeps one bf16 copy of the logits per micro-batch (T x V x 2 bytes) between forward and backward,
    # "recompute" keeps lse[T] only and recomputes the logits per vocab chunk; "auto" = the library's size policy
    bwd_mode = args.lm_head_backward
    if bwd_mode == "auto":
        bwd_mode = "saved" if P.saved_logits_policy(T, V) else "recompute"
    save_flag = bwd_mode == "saved"
    saved_blk = torch.empty(T, int(lib.prl_lmhead_saved_logits_ld(V)), dtype=torch.bfloat16, device=dev) if save_flag else None
    ws_bs = torch.empty(max(int(lib.prl_lmhead_bwd_saved_workspace_bytes(T, H, V)), 16), dtype=torch.uint8, device=dev)

    Wd = W.detach()

    def step_resident():
        """inputs already in HBM; raw C-ABI calls, no autograd bookkeeping"""
        adv_all, = (P.compute_advantages_tensor(rewards_d, lengths_d, wl["group"], "drgrpo"),)
        last = None
        for i in range(n_mb):
            mb = dev_mbs[i % len(dev_mbs)]
            hid = mb["hidden"][0]
            ids = mb["input_ids"][0]
            if save_flag:
                logp, ent, lse, _, _ = P.lm_head_forward_save_raw(hid, Wd, ids, 1.0 / mb["temperature"], workspace=ws_f, saved=saved_blk)
            else:
                logp, ent, lse, _ = P.lm_head_forward_raw(hid, Wd, ids, 1.0 / mb["temperature"], workspace=ws_f)
            loss, ratio, clipped, is_clipped, dlogp = _loss_packed_raw(
                logp, mb["logprobs"][0], mb["advantages"][0], mb["loss_mask"][0].view(torch.uint8), cu, nseg, mode,
                cfg.ratio_length_norm, cfg.clip_ratio, loss_scale)
            if save_flag:
                P.lm_head_backward_saved_raw(hid, Wd, ids, lse, dlogp, 1.0 / mb["temperature"], saved_blk, dweight=main_grad,
                                             dweight_accumulate=True, workspace=ws_bs)
            else:
                P.lm_head_backward_raw(hid, Wd, ids, lse, dlogp, 1.0 / mb["temperature"], dweight=main_grad, dweight_accumulate=True,
                                       workspace=ws_b)
            last = loss
        if dist is not None and rs_out is not None:
            dist.reduce_scatter_tensor(rs_out, main_grad.view(-1))
        return last

    copy_stream = torch.cuda.Stream(device=dev)
    h2d_bytes = sum(v.numel() * v.element_size() for v in wl["mbs"][0].values() if torch.is_tensor(v)) * n_mb
    loss_host = torch.empty(n_mb, dtype=torch.float32).pin_memory()

    def step_e2e():
        """public API (`rl_micro_step`), HOST micro-batches: H2D of every micro-batch from pinned memory
        (prefetched one ahead on a copy stream), D2H of every micro-batch's loss."""
        r_host = wl["rewards"].tolist()
        adv = P.compute_advantages(r_host, wl["lengths"].tolist(), wl["group"], "drgrpo")  # host lists in/out
        del adv
        cur = torch.cuda.current_stream()

        def upload(i):
            src = wl["mbs"][i % len(wl["mbs"])]
            with torch.cuda.stream(copy_stream):
                out = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in src.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return out, ev

        nxt = upload(0)
        for i in range(n_mb):
            mb, ev = nxt
            if i + 1 < n_mb:
                nxt = upload(i + 1)
            cur.wait_event(ev)
            for v in mb.values():
                if torch.is_tensor(v):
                    v.record_stream(cur)
            hid = mb["hidden"].requires_grad_(True)
            out = P.rl_micro_step(hid, W, mb, cfg, loss_scale, main_grad=main_grad, response_lengths=None,
                                  forward_skips_masked=args.forward_skips_masked, save_logits=save_flag)
            loss_host[i].copy_(out["loss"], non_blocking=True)
        if dist is not None and rs_out is not None:
            dist.reduce_scatter_tensor(rs_out, main_grad.view(-1))
        torch.cuda.current_stream().synchronize()
        return float(loss_host[-1])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    ms_res, _ = timed(step_resident, args.steps, args.warmup)
    launches = _lib.launch_count(reset=True)
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    ms_e2e, wall_e2e = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}

    # per-kernel-family timing (one family per launch sequence, CUDA events on the launching stream)
    fam = {}
    if rank == 0 and not args.no_breakdown:
        mb = dev_mbs[0]
        hid, ids = mb["hidden"][0], mb["input_ids"][0]
        logp, ent, lse, _ = P.lm_head_forward_raw(hid, Wd, ids, 1.0, workspace=ws_f)
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

        if save_flag:
            vc, nchunks = V, 1
            fwd_s = lambda: P.lm_head_forward_save_raw(hid, Wd, ids, 1.0, workspace=ws_f, saved=saved_blk)
            bwd_s = lambda m: P.lm_head_backward_saved_raw(hid, Wd, ids, lse, dlogp, 1.0, saved_blk, dweight=main_grad,
                                                           dweight_accumulate=True, phase_mask=m, workspace=ws_bs)
            fam["lmhead_fwd_save_kernel"] = dict(ms_per_launch=ev_time(fwd_s, 8), launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            bwd_s(1)  # the block now holds real dlogits (tensor-core power is data dependent), as inside the step
            for name, mask in (("lmhead_bwd_dhidden_kernel", 2), ("lmhead_bwd_dweight_kernel", 4)):
                fam[name] = dict(ms_per_launch=ev_time(lambda: bwd_s(mask), 4), launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            # HBM-bound in-place pass (re-applied to its own output here: timing only)
            fam["dz_from_saved_logits_kernel"] = dict(ms_per_launch=ev_time(lambda: bwd_s(1), 8), launches_per_mb=1, flops_per_launch=0.0,
                                                      bytes_per_launch=algorithmic_bytes("dz_from_saved_logits_kernel", T, H, V, V))
        else:
            vc = 16384  # default vocab chunk of prl_lmhead_bwd
            nchunks = (V + vc - 1) // vc
            fam["lmhead_fwd_kernel"] = dict(ms_per_launch=ev_time(lambda: P.lm_head_forward_raw(hid, Wd, ids, 1.0, workspace=ws_f), 8),
                                            launches_per_mb=1, flops_per_launch=2.0 * T * H * V)
            for name, mask in (("lmhead_bwd_dz_kernel", 1), ("lmhead_bwd_dhidden_kernel", 2), ("lmhead_bwd_dweight_kernel", 4)):
                t_ms = ev_time(lambda: P.lm_head_backward_raw(hid, Wd, ids, lse, dlogp, 1.0, dweight=main_grad, dweight_accumulate=True,
                                                              phase_mask=mask, workspace=ws_b), 4)
                fam[name] = dict(ms_per_launch=t_ms / nchunks, launches_per_mb=nchunks, flops_per_launch=2.0 * T * H * V / nchunks)

    peaks = measured_peaks()
    tokens_per_step = n_mb * T
    value = world * tokens_per_step * args.steps / (ms_res * 1e-3)
    e2e_value = world * tokens_per_step * args.steps / (ms_e2e * 1e-3)
    flops_per_token = 6.0 * H * V
    alg_tflops = value / world * flops_per_token / 1e12
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: H={H} V={V}, {n_mb} packed micro-batches x [1,{T}] tokens per GPU per step "
                               f"({len(wl['seg'])} x {wl['seg'][0]}-token rollouts each), ratio_type={cfg.ratio_type}, "
                               + ("all-ones mask" if args.prompt_frac == 0 else f"first {args.prompt_frac:.0%} of every rollout masked (prompt); the e2e arm's "
                                  "backward skips masked positions, the resident arm visits all T like the reference") + ", tau=1",
                   "tokens_per_gpu_per_step": tokens_per_step, "parallelism": f"token-sharded dp{world} (no data-path collective; "
                   + ("d_W_lm fp32 reduce-scatter once per step)" if world > 1 else "single GPU)"),
                   "l2_policy": "inputs larger than L2 (W_lm 1.24 GB + 134 MB hidden per micro-batch, rotating micro-batches)",
                   "flops_per_token_algorithmic": flops_per_token,
                   "lm_head_backward": (f"saved-logits: one bf16 [T, V] block ({T * V * 2 / 2**30:.2f} GiB) kept per micro-batch between "
                                        "forward and backward, rewritten in place into dlogits; 3 GEMMs of T*H*V per micro-batch"
                                        if save_flag else "recompute: lse[T] kept only, logits recomputed per 16384-column vocab "
                                        "chunk; 4 GEMMs of T*H*V per micro-batch")},
        "per_gpu": value / world,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 * n_mb + 8 * n_mb,
                "ms_per_step": ms_e2e / args.steps, "api": "prime_rl_cse_b200.rl_micro_step + compute_advantages (host buffers)"},
        "gpu_launches": launches_timed,
        "clocks": clocks,
        "tensor_tflops_algorithmic_per_gpu": alg_tflops,
        "frac_of_peak": {"measured_burst": alg_tflops / peaks["burst"], "measured_sustained": alg_tflops / peaks["sustained"],
                         "nominal_2250": alg_tflops / 2250.0, "executed_over_algorithmic": (6.0 if save_flag else 8.0) / 6.0},
    }
    if fam:
        top = max(fam, key=lambda k: fam[k]["ms_per_launch"] * fam[k]["launches_per_mb"])
        f = fam[top]
        achieved = f["flops_per_launch"] / (f["ms_per_launch"] * 1e-3) / 1e12
        result["roofline"] = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                              "frac": achieved / peaks["sustained"],
                              "traffic": NCU_TRAFFIC_BYTES[bwd_mode].get(top) if args.workload == "qwen3-8b-hendrycks" else None,
                              "traffic_source": NCU_TRAFFIC_SOURCE[bwd_mode] + " (dram read+write bytes per launch)",
                              "algorithmic_bytes": algorithmic_bytes(top, T, H, V, vc),
                              "flops_per_launch": f["flops_per_launch"], "ms_per_launch": f["ms_per_launch"],
                              "peak_source": peaks["source"] + " sustained (kernel timed inside a long power-capped step)",
                              "frac_of_burst": achieved / peaks["burst"]}
        result["kernel_breakdown_ms_per_micro_batch"] = {k: v["ms_per_launch"] * v["launches_per_mb"] for k, v in fam.items()}
        result["kernel_tflops"] = {k: v["flops_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e12 for k, v in fam.items() if v["flops_per_launch"]}
        hbm = {k: v for k, v in fam.items() if v.get("bytes_per_launch")}
        if hbm:  # the HBM-bound pass of the saved-logits backward against the measured copy bandwidth
            result["hbm_kernels"] = {k: {"achieved_gbs": v["bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9, "peak_gbs": peaks["hbm"],
                                         "frac": v["bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9 / peaks["hbm"],
                                         "algorithmic_bytes": v["bytes_per_launch"]} for k, v in hbm.items()}
    if rank == 0 and not args.no_torch_baseline:
        result["torch_gpu_reference"] = torch_gpu_reference(wl, dev)
    if rank == 0:
        if not args.no_cpu_baseline and world >= 1:
            result["cpu_baseline"] = cpu_baseline(args.workload, steps=3, warmup=0)
        emit(result)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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


def cpu_step(wl, T_s):
    from oracle import torch_port

    H, V = wl["H"], wl["V"]
    mb = wl["mbs"][0]
    nseg = 4
    seg = [T_s // nseg] * nseg
    hidden = mb["hidden"][0, :T_s].float().clone().requires_grad_(True)  # config 1 runs fp32 on the CPU
    weight = wl["weight"].float().clone().requires_grad_(True)
    ids = mb["input_ids"][0, :T_s]
    pos = torch.cat([torch.arange(n) for n in seg])
    adv_list = torch_port.advantages_cpu(wl["rewards"].tolist(), wl["lengths"].tolist(), wl["group"], "drgrpo")
    adv = torch.cat([torch.full((n,), adv_list[i]) for i, n in enumerate(seg)])
    mask = torch.ones(T_s, dtype=torch.bool)
    old = mb["logprobs"][0, :T_s]
    t0 = time.perf_counter()
    loss, _, _ = torch_port.micro_step_cpu(hidden, weight, ids, pos, old, adv, mask, 1.0, wl["ratio_type"], False, 8.0, T_s)
    return time.perf_counter() - t0, float(loss)


def cpu_baseline(workload: str, steps: int, warmup: int) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(workload, 0, n_distinct=1, pin=False)
    T_s = min(CPU_SAMPLE_TOKENS, wl["T"])
    for _ in range(warmup):
        cpu_step(wl, T_s)
    times = [cpu_step(wl, T_s)[0] for _ in range(max(steps, 1))]
    dt = statistics.median(times)
    return {"value": T_s / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one packed row of {T_s} tokens (4 segments) at H={wl['H']} V={wl['V']}, fp32, torch CPU ops "
                      f"(oracle/torch_port.py restating rl/train.py:285-310), median of {max(steps,1)} run(s), {dt:.2f} s each"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(args.workload, 0, n_distinct=1, pin=False)
    T_s = min(CPU_SAMPLE_TOKENS, wl["T"])
    for _ in range(args.warmup):
        cpu_step(wl, T_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(wl, T_s)
    dt = time.perf_counter() - t0
    value = T_s * args.steps / dt
    H, V = wl["H"], wl["V"]
    sample = (f"each step = one packed row of {T_s} tokens (4 segments) of the same workload shape, fp32, all host threads "
              f"(torch CPU ops, oracle/torch_port.py restating the reference's rl/train.py:285-310; the reference is pure Python "
              f"and /root/reference does not travel to the GPU box)")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: H={H} V={V}, bounded sample of {T_s} tokens per step", "parallelism": "host CPU, rank 0 only"},
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
    ap.add_argument("--forward-skips-masked", action="store_true", help="e2e arm: also restrict the forward to unmasked positions")
    ap.add_argument("--prompt-frac", type=float, default=0.0, help="mask the first fraction of every rollout (default 0 = all-ones mask)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the torch-on-GPU reference-chain timing")
    ap.add_argument("--lm-head-backward", default="auto", choices=["auto", "saved", "recompute"],
                    help="saved = keep the bf16 logits between forward and backward (T*V*2 bytes); recompute = keep lse only")
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
