"""L2-resident GEMM (operands 64 MB, no epilogue stores): how much does operand traffic from L2 alone cost?
Compare PRL_DEBUG_SKIPB=0/1 with PRL_DEBUG_NOSTORE=1 (both profiling aids; results are wrong by construction)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib
lib = _lib.load()
dev = "cuda"
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 8192, 2048)))
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); C = torch.zeros(M, N, device=dev)
fn = lambda: lib.prl_gemm_bf16_f32(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, 0, 0, 0, _lib.stream())
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; t0 = time.time(); e0.record()
while time.time() - t0 < 1.5:
    for _ in range(16):
        fn(); n += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"M{M} N{N} K{K} NOSTORE={os.environ.get('PRL_DEBUG_NOSTORE','0')} SKIPB={os.environ.get('PRL_DEBUG_SKIPB','0')}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TF/s", flush=True)
