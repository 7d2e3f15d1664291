"""A/B of the engine's tuning knobs at a bench shape: per-kernel-family time of one micro-batch (forward, dlogits recompute,
d_hidden, d_weight) in ~1 s sustained loops under the power cap, real operand data, CUDA events on the launching stream.

    python tools/knob_sweep.py [--shape 8b|1.7b] [--msub 0|1|2] [--only fwd,dz,dh,dw]

Knobs are set through prl_set_knob (the environment is only read once); PRL_MSUB / PRL_CTA_GROUP select kernel variants at
first use, so they are taken from the command line and set before anything runs (run the script once per value)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="8b")
ap.add_argument("--msub", type=int, default=0)
ap.add_argument("--only", default="fwd,dz,dh,dw")
ap.add_argument("--secs", type=float, default=1.0)
ap.add_argument("--hints", action="store_true", help="sweep the L2 eviction hints of the dz / fwd sweeps instead of the scheduling knobs")
a = ap.parse_args()
if a.msub:
    _lib.set_knob("PRL_MSUB", a.msub)
T, H, V = {"8b": (16384, 4096, 151936), "1.7b": (16384, 2048, 151936)}[a.shape]
dev = "cuda"
torch.manual_seed(0)
hidden = torch.randn(T, H, device=dev).bfloat16()
W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
ids = torch.randint(0, V, (T,), device=dev)
dW = torch.zeros(V, H, device=dev)
lib = _lib.load()
ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 32768), dtype=torch.uint8, device=dev)
logp, ent, lse, _ = lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)
g = torch.randn(T, device=dev) / T


def ev(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    x.record()
    while n < 3 or time.perf_counter() - t0 < a.secs:
        fn()
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    y.record()
    torch.cuda.synchronize()
    return x.elapsed_time(y) / n


FAM = {
    "fwd": lambda: lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f),
    "dz": lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=1, workspace=ws_b),
    "dh": lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=2, workspace=ws_b),
    "dw": lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, phase_mask=4, workspace=ws_b),
}
only = a.only.split(",")
lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, workspace=ws_b)  # the dlogits block holds real data

DEFAULTS = {"PRL_VCHUNK": 16384, "PRL_FWD_STAGES": 5, "PRL_DZ_STAGES": 5, "PRL_SWEEP_GROUP_M": 32, "PRL_GROUP_M": 0, "PRL_TILE_STAGES": 0}
SETTINGS = [
    ("defaults", {}, only),
    ("defaults (again)", {}, only),
    ("PRL_VCHUNK=8192", {"PRL_VCHUNK": 8192}, ["dz", "dh", "dw"]),
    ("PRL_VCHUNK=32768", {"PRL_VCHUNK": 32768}, ["dz", "dh", "dw"]),
    ("PRL_FWD_STAGES=4", {"PRL_FWD_STAGES": 4}, ["fwd"]),
    ("PRL_FWD_STAGES=6", {"PRL_FWD_STAGES": 6}, ["fwd"]),
    ("PRL_FWD_STAGES=7", {"PRL_FWD_STAGES": 7}, ["fwd"]),
    ("PRL_DZ_STAGES=4", {"PRL_DZ_STAGES": 4}, ["dz"]),
    ("PRL_DZ_STAGES=6", {"PRL_DZ_STAGES": 6}, ["dz"]),
    ("PRL_SWEEP_GROUP_M=16", {"PRL_SWEEP_GROUP_M": 16}, ["fwd", "dz"]),
    ("PRL_SWEEP_GROUP_M=64", {"PRL_SWEEP_GROUP_M": 64}, ["fwd", "dz"]),
    ("PRL_GROUP_M=4", {"PRL_GROUP_M": 4}, ["dh", "dw"]),
    ("PRL_GROUP_M=16", {"PRL_GROUP_M": 16}, ["dh", "dw"]),
    ("PRL_TILE_STAGES=2", {"PRL_TILE_STAGES": 2}, ["dh", "dw"]),
    ("defaults (end)", {}, only),
]
if os.environ.get("PRL_LIB_PATH", "").endswith("_debug.so"):
    # profiling aids of the -DPRL_DEBUG_KNOBS build (make -C prime_rl_cse_b200/csrc debug; PRL_LIB_PATH=.../libprl_b200_debug.so):
    # how much of each kernel is its epilogue's global-memory traffic (results are WRONG by design)
    DEFAULTS["PRL_DEBUG_NOSTORE"] = 0
    SETTINGS = [("defaults", {}, only), ("PRL_DEBUG_NOSTORE=1", {"PRL_DEBUG_NOSTORE": 1}, only), ("defaults (end)", {}, only)]
if a.hints:
    H3 = [(-1, -1, -1), (0, 0, 0), (2, 0, 0), (2, 1, 0), (2, 0, 1), (2, 1, 1), (0, 1, 1), (2, 2, 1), (0, 0, 1)]
    SETTINGS = [("defaults", {}, only)]
    for fam, f in (("DZ", "dz"), ("FWD", "fwd"), ("DH", "dh"), ("DW", "dw")):
        if f in only:
            SETTINGS += [(f"PRL_HINT_{fam}_A/B/C={x}", {f"PRL_HINT_{fam}_A": x[0], f"PRL_HINT_{fam}_B": x[1], f"PRL_HINT_{fam}_C": x[2]}, [f]) for x in H3[1:]]
    SETTINGS += [("defaults (end)", {}, only)]
    for fam in ("DZ", "FWD", "DH", "DW"):
        for o in "ABC":
            DEFAULTS[f"PRL_HINT_{fam}_{o}"] = -1
print(f"shape {a.shape}: T={T} H={H} V={V}  PRL_MSUB={a.msub or 'default (sweeps 1, store GEMMs 2)'}   ms per micro-batch", flush=True)
for name, kn, fams in SETTINGS:
    for k, v in DEFAULTS.items():
        _lib.set_knob(k, v)
    for k, v in kn.items():
        _lib.set_knob(k, v)
    res = {f: ev(FAM[f]) for f in fams if f in only}
    print(f"{name:24s} " + "  ".join(f"{f} {ms:7.3f}" for f, ms in res.items()) + (f"   sum {sum(res.values()):7.3f}" if len(res) == 4 else ""), flush=True)
