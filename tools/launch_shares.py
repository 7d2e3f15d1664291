"""Per-kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)."""
import csv, sys
from collections import OrderedDict

NAMES = {
    "gemm_kernel<2, 1, 0, 0, 0>": "lmhead_fwd (statistics epilogue)",
    "gemm_kernel<2, 1, 0, 0, 3>": "lmhead_fwd_save (statistics + bf16 logits store)",
    "gemm_kernel<2, 1, 0, 0, 1>": "lmhead_bwd_dz (dlogits recompute)",
    "gemm_kernel<2, 2, 0, 1, 2>": "lmhead_bwd_dhidden",
    "gemm_kernel<2, 2, 1, 1, 2>": "lmhead_bwd_dweight",
}
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in data:
    k = r[iK].replace("prl::", "").replace("void ", "")
    for pat, nice in NAMES.items():
        if pat in k:
            k = f"{pat} = {nice}"
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + float(r[iV].replace(",", "")))
tot = sum(t for _, t in agg.values())
print(f"# launches captured: {len(data)}   total device time: {tot / 1e6:.1f} ms\n")
print(f"{'kernel':75s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:75]:75s} {n:8d} {t / 1e6:10.2f} {t / n / 1e3:9.1f} {100 * t / tot:6.1f}%")
