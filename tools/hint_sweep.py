"""Sweep L2 eviction hints (PRL_HINT_<FAMILY>_<A|B|C>) per kernel family at the Qwen3-8B shape, ~1 s sustained each."""
import itertools, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw

lib = _lib.load()
dev = "cuda"
T, H, V = 16384, 4096, 151936
torch.manual_seed(0)
hidden = torch.randn(T, H, device=dev).bfloat16()
W = (torch.randn(V, H, device=dev) * 0.02).bfloat16()
ids = torch.randint(0, V, (T,), device=dev)
dW = torch.zeros(V, H, device=dev)
g = torch.randn(T, device=dev) / T
ws_f = torch.empty(lib.prl_lmhead_fwd_workspace_bytes(T, H, V), dtype=torch.uint8, device=dev)
ws_b = torch.empty(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, 32768), dtype=torch.uint8, device=dev)
logp, ent, lse, _ = lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)
# fill the dlogits block with REAL data first: tensor-core power is data dependent (an all-zero operand runs ~25 % faster)
lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, workspace=ws_b, phase_mask=1)
torch.cuda.synchronize()
FL = 2.0 * T * H * V


_saved = {}


def saved_block():
    """[T, ld] bf16 block of the saved-logits mode, holding REAL dlogits (forward, then the in-place conversion)"""
    if "blk" not in _saved:
        from prime_rl_cse_b200 import lm_head_backward_saved_raw, lm_head_forward_save_raw

        _, _, _, _, blk = lm_head_forward_save_raw(hidden, W, ids, 1.0, workspace=ws_f)
        lm_head_backward_saved_raw(hidden, W, ids, lse, g, 1.0, blk, want_dhidden=False, want_dweight=False, phase_mask=1)
        torch.cuda.synchronize()
        _saved["blk"] = blk
    return _saved["blk"]


def fam_fn(fam):
    if fam == "FWD":
        return lambda: lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f)
    if fam in ("FWDS", "SDZ", "SDH", "SDW", "SBWD", "STEP_S"):
        # saved-logits mode.  SDZ re-applies the in-place conversion to its own output (timing only)
        from prime_rl_cse_b200 import lm_head_backward_saved_raw, lm_head_forward_save_raw

        blk = saved_block()
        fwd = lambda: lm_head_forward_save_raw(hidden, W, ids, 1.0, workspace=ws_f, saved=blk)
        bwd = lambda m: lm_head_backward_saved_raw(hidden, W, ids, lse, g, 1.0, blk, dweight=dW, dweight_accumulate=True, phase_mask=m)
        if fam == "FWDS":
            return fwd
        if fam == "STEP_S":
            return lambda: (fwd(), bwd(7))
        return lambda: bwd({"SDZ": 1, "SDH": 2, "SDW": 4, "SBWD": 7}[fam])
    if fam == "STEP_R":
        return lambda: (lm_head_forward_raw(hidden, W, ids, 1.0, workspace=ws_f),
                        lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, workspace=ws_b))
    mask = {"DZ": 1, "DH": 2, "DW": 4, "BWD": 7, "DZDH": 3, "DZDW": 5, "DHDW": 6}[fam]
    return lambda: lm_head_backward_raw(hidden, W, ids, lse, g, 1.0, dweight=dW, dweight_accumulate=True, workspace=ws_b, phase_mask=mask)


def run(fam, hints, secs=1.0):
    for k, v in zip("ABC", hints):
        os.environ[f"PRL_HINT_{fam}_{k}"] = str(v)
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
    print(f"{fam} hints A,B,C={hints}: {ms:.2f} ms  {FL/ms/1e9:.0f} TF/s", flush=True)
    for k in "ABC":
        os.environ.pop(f"PRL_HINT_{fam}_{k}", None)
    return ms


if __name__ == "__main__":
    fams = sys.argv[1:] or ["DH", "DW", "DZ", "FWD"]
    for fam in fams:
        best = None
        combos = itertools.product((0, 1, 2), (0, 1, 2), (0, 1)) if fam != "FWD" else itertools.product((0, 1, 2), (0, 1, 2), (0,))
        for h in combos:
            ms = run(fam, h)
            if best is None or ms < best[0]:
                best = (ms, h)
        print(f"== best {fam}: {best}", flush=True)
