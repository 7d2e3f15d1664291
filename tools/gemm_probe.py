"""Bring-up probe for the tcgen05 GEMM engine: per-majorness correctness with error structure, then
quick timings (plain GEMM, fused forward, backward) with CUDA events.  Prints; asserts nothing."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw  # noqa: E402

lib = _lib.load()
dev = "cuda"


def gemm(A, B, a_mn, b_mn, M, N, K):
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    C = torch.full((M, N), float("nan"), device=dev)
    rc = lib.prl_gemm_bf16_f32(_lib.ptr(a_st), _lib.ptr(b_st), _lib.ptr(C), M, N, K, a_mn, b_mn, 0, _lib.stream())
    _lib.check(rc, "gemm")
    torch.cuda.synchronize()
    return C


def correctness():
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 128), (304, 520, 200)]:
        torch.manual_seed(0)
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        want = A.float() @ B.float().t()
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                C = gemm(A, B, a_mn, b_mn, M, N, K)
                err = (C - want).abs()
                bad = err > 1e-2 * want.abs().max()
                nan = torch.isnan(C).sum().item()
                msg = f"M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: max_err={err[~torch.isnan(err)].max().item() if nan < C.numel() else float('nan'):.4g} bad={bad.sum().item()}/{C.numel()} nan={nan}"
                if bad.any():
                    rows = bad.any(dim=1).nonzero().flatten()[:8].tolist()
                    cols = bad.any(dim=0).nonzero().flatten()[:8].tolist()
                    msg += f" first bad rows {rows} cols {cols}"
                print(msg, flush=True)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def perf():
    M = N = K = 8192
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev)
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 1)]:
        a_st = A.t().contiguous() if a_mn else A
        b_st = B.t().contiguous() if b_mn else B
        ms = timeit(lambda: lib.prl_gemm_bf16_f32(_lib.ptr(a_st), _lib.ptr(b_st), _lib.ptr(C), M, N, K, a_mn, b_mn, 0, _lib.stream()))
        print(f"gemm 8192^3 a_mn={a_mn} b_mn={b_mn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    ms = timeit(lambda: torch.matmul(A, B.t()))
    print(f"torch.matmul bf16 8192^3: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

    for (T, H, V) in [(4096, 1024, 151936), (16384, 2048, 151936), (16384, 4096, 151936)]:
        hidden = torch.randn(T, H, device=dev).bfloat16()
        W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
        ids = torch.randint(0, V, (T,), device=dev)
        out = {}

        def fwd():
            out["r"] = lm_head_forward_raw(hidden, W, ids, 1.0)

        ms_f = timeit(fwd, iters=3, warm=1)
        logp, ent, lse, _ = out["r"]
        g = torch.randn(T, device=dev) / T
        dW = torch.zeros(V, H, device=dev)

        def bwd():
            lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True)

        ms_b = timeit(bwd, iters=2, warm=1)
        fl = 2.0 * T * H * V
        print(f"T{T} H{H} V{V}: fwd {ms_f:.2f} ms ({fl/ms_f/1e9:.0f} TF/s exec)  bwd {ms_b:.2f} ms ({3*fl/ms_b/1e9:.0f} TF/s exec)  "
              f"fwd+bwd {T/(ms_f+ms_b)*1e3:.0f} tok/s  alg {3*fl/(ms_f+ms_b)/1e9:.0f} TF/s  ent_mean {ent.mean().item():.3f}", flush=True)
        del hidden, W, dW


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    correctness()
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        perf()
