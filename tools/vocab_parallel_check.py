"""Vocab-sharded LM head (BASELINE.json config 4) on real GPUs: run under torchrun, one rank per GPU.
Every rank builds the same inputs, keeps its FSDP-style vocab shard of W_lm, runs the sharded fused
forward/backward (partials exchanged over NCCL) and compares with the unsharded fused path on its own GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/vocab_parallel_check.py [--perf]
"""
import os, sys, time
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prime_rl_cse_b200 as P


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    perf = "--perf" in sys.argv
    T, H, V, tau = (16384, 5120, 151936, 1.0) if perf else (2048, 5120, 151936, 0.6)
    g = torch.Generator(device="cuda").manual_seed(1234)  # same on every rank
    hidden = torch.randn(T, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * 0.02).bfloat16()
    ids = torch.randint(0, V, (T,), device="cuda", generator=g)
    gl = torch.randn(T, device="cuda", generator=g) / T
    v0, v1 = P.vocab_shard_bounds(V, world, rank)
    Ws = W[v0:v1].clone().requires_grad_(True)

    h = hidden.clone().requires_grad_(True)
    lp, ent = P.vocab_parallel_lm_head_logprobs_entropy(h.unsqueeze(0), Ws, ids.unsqueeze(0), tau, V)
    (lp[0] * gl).sum().backward()

    h2 = hidden.clone().requires_grad_(True)
    W2 = W.clone().requires_grad_(True)
    lp2, ent2 = P.fused_lm_head_logprobs_entropy(h2.unsqueeze(0), W2, ids.unsqueeze(0), tau)
    (lp2[0] * gl).sum().backward()
    torch.cuda.synchronize()

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()

    res = dict(dlogp=(lp - lp2).abs().max().item(), dent=(ent - ent2).abs().max().item(), dh=rel(h.grad, h2.grad),
               dw=rel(Ws.grad, W2.grad[v0:v1]))
    ok = res["dlogp"] < 5e-4 and res["dent"] < 5e-4 and res["dh"] < 5e-3 and res["dw"] < 1e-3
    print(f"rank {rank}/{world} shard [{v0},{v1}): {res} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)

    # on-device logging statistics across ranks (row N1): every rank logs different values; mean / median / min / max must be
    # those of the union (the reference all-gathers every masked token, trainer/utils.py:241-252)
    gs = torch.Generator(device="cuda").manual_seed(77 + rank)
    vals = torch.randn(5000 + 1000 * rank, device="cuda", generator=gs).round(decimals=2)
    msk = torch.rand(vals.numel(), device="cuda", generator=gs) < 0.5
    st = P.TensorStats()
    st.add("x", vals[:3000], msk[:3000])
    st.add("x", vals[3000:], msk[3000:])
    got = st.compute_stats()
    mine = vals[msk]
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], device="cuda"))
    padded = torch.zeros(max(int(x) for x in sizes), device="cuda")
    padded[: mine.numel()] = mine
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    union = torch.cat([b[: int(n)] for b, n in zip(bufs, sizes)])
    ok_s = (got["x/median"] == torch.median(union).item() and got["x/count"] == union.numel() and got["x/min"] == union.min().item()
            and got["x/max"] == union.max().item() and abs(got["x/mean"] - union.double().mean().item()) < 1e-9)
    print(f"rank {rank}: TensorStats over {union.numel()} values of {world} ranks: median {got['x/median']} -> {'OK' if ok_s else 'MISMATCH'}", flush=True)
    flag2 = torch.tensor([0 if ok_s else 1], device="cuda")
    dist.all_reduce(flag2)
    flag += flag2

    if perf:
        def step():
            hh = hidden.detach().requires_grad_(True)
            l, e = P.vocab_parallel_lm_head_logprobs_entropy(hh.unsqueeze(0), Ws, ids.unsqueeze(0), tau, V)
            (l[0] * gl).sum().backward()
        for _ in range(2):
            step()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            step()
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"vocab-sharded x{world} Qwen3-32B shape T={T}: {ms.item():.2f} ms per micro-batch -> {T / ms.item() * 1e3:.0f} tok/s "
                  f"(all ranks work on the same tokens), alg {6.0 * H * V * T / ms.item() / 1e9 / world:.0f} TFLOP/s per GPU", flush=True)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
