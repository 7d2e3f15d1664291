"""Sustained (power-capped) throughput of the GEMM engine per operand majorness / epilogue / shape, with NVML
clock + power sampling at 10 ms.  Diagnostic only."""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib
import pynvml

lib = _lib.load()
dev = "cuda"
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler:
    def __init__(self):
        self.run = True; self.clk = []; self.pw = []
        self.t = threading.Thread(target=self.loop, daemon=True); self.t.start()
    def loop(self):
        while self.run:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.01)
    def stop(self):
        self.run = False; self.t.join()
        n = len(self.clk) // 2  # second half = steady state
        c = self.clk[n:]; p = self.pw[n:]
        return sum(c) / len(c), sum(p) / len(p)


def sustained(fn, flops, secs=1.2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = Sampler()
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(8):
            fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    clk, pw = s.stop()
    ms = e0.elapsed_time(e1) / n
    return flops / ms / 1e9, clk, pw, ms


def gemm_case(name, M, N, K, a_mn, b_mn, acc):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    a_st = A.t().contiguous() if a_mn else A; b_st = B.t().contiguous() if b_mn else B
    C = torch.zeros(M, N, device=dev)
    tf, clk, pw, ms = sustained(lambda: lib.prl_gemm_bf16_f32(_lib.ptr(a_st), _lib.ptr(b_st), _lib.ptr(C), M, N, K, a_mn, b_mn, acc, _lib.stream()), 2.0 * M * N * K)
    print(f"{name:34s} M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn} acc={acc}: {ms:.3f} ms {tf:7.1f} TF/s  clk {clk:6.0f} MHz  {pw:5.0f} W  -> {tf/(clk*148*8192/1e6)*100:.1f}% of clock peak", flush=True)


def lmhead_cases():
    from prime_rl_cse_b200 import lm_head_backward_raw, lm_head_forward_raw
    T, H, V = 16384, 4096, 32768
    hidden = torch.randn(T, H, device=dev).bfloat16()
    W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
    ids = torch.randint(0, V, (T,), device=dev)
    dW = torch.zeros(V, H, device=dev)
    ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
    ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 0), dtype=torch.uint8, device=dev)
    logp, ent, lse, _ = lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)
    g = torch.randn(T, device=dev) / T
    fl = 2.0 * T * H * V
    for name, fn in [("lmhead fwd (stats epilogue)", lambda: lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)),
                     ("lmhead bwd dz", lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=1, workspace=ws_b)),
                     ("lmhead bwd dh", lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=2, workspace=ws_b)),
                     ("lmhead bwd dw", lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=4, workspace=ws_b))]:
        tf, clk, pw, ms = sustained(fn, fl)
        print(f"{name:34s} T{T} H{H} V{V}: {ms:.3f} ms {tf:7.1f} TF/s  clk {clk:6.0f} MHz  {pw:5.0f} W  -> {tf/(clk*148*8192/1e6)*100:.1f}% of clock peak", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which != "all":
        if which == "lmhead":
            lmhead_cases()
        elif which == "square":
            gemm_case("square kk store", 8192, 8192, 8192, 0, 0, 0)
        sys.exit(0)
    M = N = K = 8192
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    tf, clk, pw, ms = sustained(lambda: torch.matmul(A, B.t()), 2.0 * M * N * K)
    print(f"{'cublas 8192^3':34s}: {ms:.3f} ms {tf:7.1f} TF/s  clk {clk:6.0f} MHz {pw:5.0f} W -> {tf/(clk*148*8192/1e6)*100:.1f}% of clock peak", flush=True)
    del A, B
    gemm_case("square kk store", 8192, 8192, 8192, 0, 0, 0)
    gemm_case("square mm store", 8192, 8192, 8192, 1, 1, 0)
    gemm_case("fwd-like kk store (logits)", 16384, 8192, 4096, 0, 0, 0)
    gemm_case("dh-like km store", 16384, 4096, 8192, 0, 1, 0)
    gemm_case("dh-like km rmw", 16384, 4096, 8192, 0, 1, 1)
    gemm_case("dh-like kk rmw", 16384, 4096, 8192, 0, 0, 1)
    gemm_case("dw-like mm store", 8192, 4096, 16384, 1, 1, 0)
    gemm_case("dw-like mm rmw", 8192, 4096, 16384, 1, 1, 1)
    gemm_case("dw-like kk rmw", 8192, 4096, 16384, 0, 0, 1)


