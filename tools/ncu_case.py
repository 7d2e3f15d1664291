"""Short single-GPU case for ncu: one fused forward + one backward at the Qwen3-8B tile shape but a
reduced vocabulary (2 vocab chunks), so every launch has the same per-launch work as in bench.py.
NCU_SAVED=1: the saved-logits pair instead (whole-vocabulary launches: set NCU_V=151936 for the bench.py shape).
NCU_SCATTER=1: the recompute backward whose d_weight GEMM epilogue delivers (local fp32 accumulator + its tiles) into the shards
of 8 owners (prl_lmhead_bwd_dw_scatter, the default d_W_lm sync of bench.py at N > 1), the owners emulated as separate buffers on
this GPU: per-launch work as in the multi-GPU run, without the NVLink hop."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
from types import SimpleNamespace
from prime_rl_cse_b200 import (lm_head_backward_dw_scatter_raw, lm_head_backward_raw, lm_head_backward_saved_raw, lm_head_forward_raw,
                               lm_head_forward_save_raw)

T, H, V = 16384, 4096, int(os.environ.get("NCU_V", 16384))
dev = "cuda"
torch.manual_seed(0)
hidden = torch.randn(T, H, device=dev).bfloat16()
W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
ids = torch.randint(0, V, (T,), device=dev)
dW = torch.zeros(V, H, device=dev)
saved = None
acc = None
if os.environ.get("NCU_SCATTER"):
    peers = int(os.environ.get("NCU_PEERS", 8))
    per = -(-V // peers)
    shards = [torch.zeros(max(min(per, V - j * per), 0), H, device=dev) for j in range(peers)]
    ptrs = (ctypes.c_void_p * peers)(*[s_.data_ptr() for s_ in shards])
    acc = SimpleNamespace(V=V, H=H, world=peers, rank=3, rows_per_peer=per, buf=SimpleNamespace(ptrs=ptrs))
    dW.normal_()  # stands for the local total of the first n - 1 micro-batches
for _ in range(int(os.environ.get("NCU_ITERS", 2))):
    if os.environ.get("NCU_SAVED"):
        logp, ent, lse, _, saved = lm_head_forward_save_raw(hidden, W, ids, 1.0, saved=saved)
        g = torch.randn(T, device=dev) / T
        lm_head_backward_saved_raw(hidden, W, ids, lse, g, 1.0, saved, dweight=dW, dweight_accumulate=True)
        continue
    logp, ent, lse, _ = lm_head_forward_raw(hidden, W, ids, 1.0)
    g = torch.randn(T, device=dev) / T
    if os.environ.get("NCU_FWD_ONLY"):
        continue
    if acc is not None:
        lm_head_backward_dw_scatter_raw(hidden, W, ids, lse, g, 1.0, acc, local_add=dW)
        continue
    lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True)
torch.cuda.synchronize()
print("ok", float(logp.mean()), float(ent.mean()))
