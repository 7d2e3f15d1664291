"""Cost of MN-major operands: the plain GEMM entry (prl_gemm_bf16_f32, 512x256 tiles, fp32 TMA-store epilogue) in all four
operand major-ness combinations at the shape of one vocab chunk of the gradient GEMMs, ~1 s sustained loops, real data."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib

lib = _lib.load()
M, N, K = 16384, 4096, 16384
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
At, Bt = A.t().contiguous(), B.t().contiguous()
C = torch.empty(M, N, device="cuda")


def ev(fn, secs=1.0):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    x.record()
    while n < 3 or time.perf_counter() - t0 < secs:
        fn(); n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    y.record(); torch.cuda.synchronize()
    return x.elapsed_time(y) / n


for rnd in range(2):
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
        a, b = (At if a_mn else A), (Bt if b_mn else B)
        ms = ev(lambda: _lib.check(lib.prl_gemm_bf16_f32(_lib.ptr(a), _lib.ptr(b), _lib.ptr(C), M, N, K, a_mn, b_mn, 0, _lib.stream()), "gemm"))
        print(f"round {rnd}: A {'MN' if a_mn else 'K '}-major, B {'MN' if b_mn else 'K '}-major: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
ms = ev(lambda: torch.matmul(A, B.t()))
print(f"cuBLAS (torch.matmul, bf16 out): {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s")
