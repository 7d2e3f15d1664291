"""Sweep scheduling knobs (read from the environment at every launch) on the Qwen3-8B-shaped micro-step,
sustained ~1.5 s per setting; prints tokens/s.  Diagnostic only."""
import itertools, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw

lib = _lib.load()
dev = "cuda"
T, H, V = 16384, int(os.environ.get("TUNE_H", 4096)), 151936
torch.manual_seed(0)
hidden = [torch.randn(T, H, device=dev).bfloat16() for _ in range(2)]
W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
ids = torch.randint(0, V, (T,), device=dev)
dW = torch.zeros(V, H, device=dev)
g = torch.randn(T, device=dev) / T
ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 32768), dtype=torch.uint8, device=dev)


def step(i, phases):
    h = hidden[i % 2]
    logp, ent, lse, _ = lm_head_forward_raw(h, W, ids, 1.0, workspace=ws_f)
    if phases:
        lm_head_backward_raw(h, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, workspace=ws_b, phase_mask=phases)


def run(env, secs=1.5, phases=7):
    for k, v in env.items():
        os.environ[k] = str(v)
    step(0, phases); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(4):
            step(n, phases); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{env}: {ms:.2f} ms/micro-batch  {T/ms*1e3:.0f} tok/s", flush=True)
    for k in env:
        os.environ.pop(k, None)


if __name__ == "__main__":
    run({})
    for v in (8, 16, 64, 128):
        run({"PRL_SWEEP_GROUP_M": v})
    for v in (2, 4, 16, 32, 64):
        run({"PRL_GROUP_M": v})
    for v in (4096, 16384, 32768):
        run({"PRL_VCHUNK": v})
    run({})
