"""Knob sweep for the saved-logits kernels at the Qwen3-8B shape (~1 s sustained per point): scheduling band of the
whole-vocabulary d_hidden / d_weight GEMMs (PRL_GROUP_M), their L2 hints, and the forward's store hint / prefetch depth."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from hint_sweep import FL, fam_fn  # noqa: E402


def measure(fam, env, secs=1.0):
    for k, v in env.items():
        os.environ[k] = str(v)
    fn = fam_fn(fam)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(4):
            fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    for k in env:
        os.environ.pop(k, None)
    print(f"{fam:5s} {env}: {ms:.2f} ms  {FL / ms / 1e9:.0f} TF/s", flush=True)
    return ms


what = sys.argv[1:] or ["band", "hints", "fwd"]
if "band" in what:
    for fam in ("SDH", "SDW"):
        for gm in (4, 8, 16, 32, 64):
            measure(fam, {"PRL_GROUP_M": gm})
if "hints" in what:
    for fam, key in (("SDH", "DH"), ("SDW", "DW")):
        for a, b in itertools.product((0, 1, 2), (0, 1, 2)):
            measure(fam, {f"PRL_HINT_{key}_A": a, f"PRL_HINT_{key}_B": b})
if "fwd" in what:
    for st, c in itertools.product((4, 5, 6), (0, 1)):
        measure("FWDS", {"PRL_FWD_STAGES": st, "PRL_HINT_FWDS_C": c})
if "ab" in what:  # interleaved A/B of the candidates (order / temperature effects are ~1-2 %)
    for _ in range(3):
        for gm in (4, 8):
            measure("SDH", {"PRL_GROUP_M": gm})
        for gm in (4, 8):
            measure("SDW", {"PRL_GROUP_M": gm})
        for st in (4, 5):
            measure("FWDS", {"PRL_FWD_STAGES": st})
        for st in (4, 5):
            measure("FWD", {"PRL_FWD_STAGES": st})
if "dwc" in what:  # L2 policy of the fp32 reduce-add into main_grad (d_weight GEMM output), interleaved
    for _ in range(3):
        for c in (0, 1, 2):
            measure("SDW", {"PRL_HINT_DW_C": c})
