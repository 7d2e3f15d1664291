"""Achieved HBM GB/s of the HBM-bound kernels (advantages, packed loss, segments, logits statistics) at sizes
far beyond L2, against the measured copy bandwidth in MEASURED_PEAKS.json.  CUDA events, 3 warm-ups, inputs > L2."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prime_rl_cse_b200 as P
from prime_rl_cse_b200 import _lib
from prime_rl_cse_b200.loss import _loss_packed_raw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6416.7
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = json.load(open(pp))["hbm_gbs"]
dev = "cuda"


QUICK = bool(os.environ.get("HBM_QUICK"))  # under ncu: one launch per kernel is enough (ncu replays it itself)


def ev(fn, reps=10):
    if QUICK:
        reps = 1
    for _ in range(0 if QUICK else 3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    print(f"{name:44s} {nbytes/1e6:9.1f} MB  {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of measured {peak:.0f} GB/s  {note}", flush=True)


torch.manual_seed(0)
# advantages: 16 B per rollout (4 reward + 8 length (opo) + 4 out); drgrpo reads no lengths: 8 B per rollout
for G in (16, 32):
    n = 1 << 26
    r = torch.rand(n, device=dev); ln = torch.randint(1, 4096, (n,), device=dev)
    report(f"advantages drgrpo G={G} n=2^26", 8 * n, ev(lambda: P.compute_advantages_tensor(r, ln, G, "drgrpo")), "8 B/rollout")
    report(f"advantages opo    G={G} n=2^26", 16 * n, ev(lambda: P.compute_advantages_tensor(r, ln, G, "opo")), "16 B/rollout")
    del r, ln
# small (BASELINE config 3: 1024 rollouts) -> launch-latency bound
r = torch.rand(1024, device=dev); ln = torch.randint(1, 4096, (1024,), device=dev)
report("advantages drgrpo G=16 n=1024 (cfg 3)", 8 * 1024, ev(lambda: P.compute_advantages_tensor(r, ln, 16, "drgrpo"), 50), "latency-bound")

# packed loss, token mode: 13 B read + 16 B written per token
T = 1 << 26
logp = -torch.rand(T, device=dev); old = logp + 0.1 * torch.randn(T, device=dev); adv = torch.randn(T, device=dev)
mask = (torch.rand(T, device=dev) < 0.8).view(torch.uint8)
cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
report("loss token mode T=2^26", 29 * T, ev(lambda: _loss_packed_raw(logp, old, adv, mask, cu, 1, 0, False, 8.0, T)), "29 B/token")
T2 = 16384
report("loss token mode T=16384 (one micro-batch)", 29 * T2, ev(lambda: _loss_packed_raw(logp[:T2], old[:T2], adv[:T2], mask[:T2], torch.tensor([0, T2], dtype=torch.int32, device=dev), 1, 0, False, 8.0, T2), 50), "latency-bound")
nseg = 2048
cu2 = torch.arange(0, nseg + 1, dtype=torch.int32, device=dev) * (T // nseg)
report("loss sequence mode T=2^26, 2048 segments", 17 * T, ev(lambda: _loss_packed_raw(logp, old, adv, mask, cu2, nseg, 1, True, 8.0, nseg)), "13 B read + 4 B written per token")
del logp, old, adv, mask

# masked rows (rows.py): partition of the positions (latency-bound, one CTA), gather of hidden rows into "active first" order and
# the scatter of d_hidden back: 2 x 2 B x T x H each
from prime_rl_cse_b200 import rows as R_
Tm, Hm = 16384, 4096
hid = torch.randn(Tm, Hm, device=dev).bfloat16(); idsm = torch.randint(0, 1000, (Tm,), device=dev)
mk = torch.rand(Tm, device=dev) < 0.75
part = R_.partition_positions(mk, True)
na = part.n_active
report("partition_positions T=16384 (single CTA)", 1 * Tm + 8 * Tm, ev(lambda: R_.partition_positions(mk, True), 50), "latency-bound by design")
report("gather_rows 16383 x 4096 bf16", 2 * 2 * (Tm - 1) * Hm, ev(lambda: R_.gather_rows(hid, idsm, part, part.rows)), "2 B read + 2 B written per element")
hp, _ = R_.gather_rows(hid, idsm, part, na)
report("scatter_rows (75 % active) -> 16384 x 4096", 2 * na * Hm + 2 * Tm * Hm, ev(lambda: R_.scatter_rows(hp, part, na, Hm)), "active rows read, all rows written")
del hid, hp

# segments: 8 B/token, single CTA by design (T <= 64k in the trainer)
pos = torch.arange(32768, device=dev) % 4096
report("segments T=32768 (single CTA)", 8 * 32768, ev(lambda: P.segments_from_position_ids(pos), 50), "latency-bound by design")

# logits statistics on materialised fp32 logits: 4 B/logit
R, V = 4096, 151936
x = torch.randn(1, R, V, device=dev); ids = torch.randint(0, V, (1, R), device=dev)
report("logits stats (logp+entropy) 4096 x 151936", 4 * R * V, ev(lambda: P.shifted_logprobs_entropy(x, ids, 1.0), 5), "4 B/logit")
