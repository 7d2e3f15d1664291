"""Sustained forward / dlogits throughput vs token count and hidden size (is the A operand's L2 residency the limiter?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw
lib = _lib.load(); dev = "cuda"
V = 151936
for (T, H) in [(16384, 4096), (8192, 4096), (4096, 4096), (16384, 2048), (32768, 2048), (32768, 4096)]:
    torch.manual_seed(0)
    hidden = torch.randn(T, H, device=dev).bfloat16(); W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
    ids = torch.randint(0, V, (T,), device=dev); g = torch.randn(T, device=dev) / T
    ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
    ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 0), dtype=torch.uint8, device=dev)
    logp, ent, lse, _ = lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)
    for name, fn in (("FWD", lambda: lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)),
                     ("DZ ", lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, want_dweight=False, phase_mask=1, workspace=ws_b))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0; t0 = time.time(); e0.record()
        while time.time() - t0 < 1.2:
            for _ in range(4):
                fn(); n += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"T={T:6d} H={H}: {name} {ms:7.2f} ms  {2.0*T*H*V/ms/1e9:6.0f} TF/s  ({ms/T*16384:6.2f} ms per 16384 tokens)", flush=True)
    del hidden, W, ws_f, ws_b
