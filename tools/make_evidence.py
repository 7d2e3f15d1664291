"""Regenerates the per-kernel evidence under profiles/ from the built library and the ncu reports a GPU run left in gpurun_out/
(nothing here is typed by hand; bench.py reads `roofline.traffic` from the JSON this writes).

    python tools/make_evidence.py --tag r2 [--reports gpurun_out/r2_*.ncu-rep]

  profiles/<tag>_sass_mnemonics.txt      per-kernel SASS mnemonic counts of libprl_b200.so (tools/sass_mnemonics.py)
  profiles/<tag>_sass/<kernel>.sass.gz   full SASS listing (gzip: 100-400 KB of text each) of every kernel on the hot path (the GEMM instantiations the defaults
                                         select, the saved-logits pass, loss / advantage / segment / row kernels)
  profiles/<tag>_ncu_<report>.txt        one line per captured launch: duration, tensor-pipe activity, DRAM bytes, L2 hit rate, ...
  profiles/<tag>_traffic.json            {workload: {mode: {kernel: {dram_bytes_per_launch, achieved_gbs, source}}}} per launch
"""
import argparse
import csv
import glob
import gzip
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "prime_rl_cse_b200", "libprl_b200.so")

# gemm_kernel<CG, MS, A_MN, B_MN, EPI> instantiations the defaults launch -> the name the host code gives the launch
GEMM_NAMES = {
    (2, 1, 0, 0, 0): "lmhead_fwd_kernel",
    (2, 1, 0, 0, 3): "lmhead_fwd_save_kernel",
    (2, 1, 0, 0, 1): "lmhead_bwd_dz_kernel",
    (2, 2, 0, 1, 2): "lmhead_bwd_dhidden_kernel",
    (2, 2, 1, 1, 2): "lmhead_bwd_dweight_kernel",
}
HOT_PLAIN = ["advantages_vec8_kernel", "dz_from_saved_logits_kernel", "merge_partials_kernel", "f32_to_bf16_kernel", "loss_token_kernel", "loss_seq_kernel",
             "loss_finalize_kernel", "advantages_small_group_kernel", "advantages_kernel", "segments_kernel", "partition_positions_kernel",
             "gather_rows_kernel", "scatter_rows_kernel", "gather_tokens_kernel", "scatter_tokens_kernel", "logits_stats_kernel",
             "logits_logsoftmax_bwd_kernel", "masked_moments_kernel", "pack_rollouts_kernel"]


def gemm_key(name: str):
    m = re.search(r"gemm_kernelILi(\d)ELi(\d)ELb(\d)ELb(\d)ELi(\d)E", name) or re.search(
        r"gemm_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(bool\))?(\d|true|false), (?:\(bool\))?(\d|true|false), (?:\(int\))?(\d)>", name)
    if not m:
        return None
    conv = lambda x: {"true": 1, "false": 0}.get(x, x)  # noqa: E731
    return tuple(int(conv(g)) for g in m.groups())


def friendly(name: str):
    k = gemm_key(name)
    if k is not None:
        return GEMM_NAMES.get(k, "gemm_kernel<%d,%d,%d,%d,%d>" % k)
    m = re.search(r"_ZN3prl\d+(\w+?_kernel)", name) or re.search(r"^(?:void )?(?:prl::)?(\w+)\s*[(<]", name)
    return m.group(1) if m else name


def _gz_text(path):
    """text file inside a gzip member without a timestamp or file name, so that regenerating identical content gives identical bytes"""
    raw = open(path, "wb")
    return io.TextIOWrapper(gzip.GzipFile(filename="", fileobj=raw, mode="wb", compresslevel=9, mtime=0), encoding="utf-8")


def write_sass(tag):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    chunks = re.split(r"(?=\t\tFunction : )", out)
    d = os.path.join(ROOT, "profiles", f"{tag}_sass")
    os.makedirs(d, exist_ok=True)
    written = []
    for ch in chunks:
        m = re.match(r"\t\tFunction : (\S+)", ch)
        if not m:
            continue
        mangled = m.group(1)
        k = gemm_key(mangled)
        if k is not None:
            if k not in GEMM_NAMES:
                continue
            short = f"{GEMM_NAMES[k]}__gemm_kernel_{k[0]}_{k[1]}_{k[2]}_{k[3]}_{k[4]}"
        else:
            base = friendly(mangled)
            if base not in HOT_PLAIN:
                continue
            t = re.search(r"_kernelILi(\d+)E", mangled)  # a template instantiation: keep its argument in the file name
            short = base + (f"_{t.group(1)}" if t else "")
        # keep address + instruction, drop the hex encodings (they double the size and carry nothing a reader needs)
        body = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln) for ln in ch.splitlines()]
        body = [ln for ln in body if ln.strip()]
        with _gz_text(os.path.join(d, short + ".sass.gz")) as f:
            f.write(f"// cuobjdump -sass prime_rl_cse_b200/libprl_b200.so, function {mangled} (instruction encodings stripped)\n" + "\n".join(body) + "\n")
        written.append(short)
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_mnemonics.txt"), "w") as f:
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_mnemonics.py"), LIB], capture_output=True, text=True).stdout)
    print(f"SASS: {len(written)} listings under profiles/{tag}_sass/, mnemonic counts in profiles/{tag}_sass_mnemonics.txt")


WANT = [("gpu__time_duration.sum", "dur"), ("sm__cycles_elapsed.avg.per_second", "sm_clk"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%active"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%elapsed"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__cluster_size", "cluster"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ns(v, unit):
    x = float(v.replace(",", ""))
    return x * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)


def summarise(rep, tag, traffic):
    """`rep`: an .ncu-rep, or the CSV `ncu -i <rep> --page raw --csv` printed on the GPU box (the reports themselves are 20-30 MB
    each and do not fit gpurun's 64 MiB return channel)"""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(f"{rep}: no launches")
        return
    hdr, units, data = rows[0], rows[1], rows[2:]
    iname = hdr.index("Kernel Name")
    idx = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    base = os.path.splitext(os.path.basename(rep))[0]
    name = base if base.startswith(tag) else f"{tag}_{base}"
    lines = ["# " + " | ".join(["kernel"] + [f"{n}[{units[i]}]" for i, n in idx]), f"# source: ncu --set full --clock-control none, report {os.path.basename(rep)}"]
    # workload / mode are encoded in the report name: <tag>_<workload>__<mode>.ncu-rep; anything else goes under "misc"
    m = re.match(rf"{tag}_(.+)__(recompute|saved|hbm|scatter)$", base)
    workload, mode = (m.group(1), m.group(2)) if m else ("misc", "misc")
    for r in data:
        fn = friendly(r[iname])
        lines.append(" | ".join([fn] + [r[i] for i, _ in idx]))
        try:
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            ns = to_ns(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
        except (ValueError, KeyError):
            continue
        rec = traffic.setdefault(workload, {}).setdefault(mode, {}).setdefault(fn, {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
        rec["launches"] += 1
        rec["dram_bytes"] += rd + wr
        rec["ns"] += ns
        rec["source"] = f"profiles/{name}.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches)"
    with open(os.path.join(ROOT, "profiles", name + ".txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{rep}: {len(data)} launches -> profiles/{name}.txt")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--reports", nargs="*", default=None)
    ap.add_argument("--no-sass", action="store_true")
    a = ap.parse_args()
    if not a.no_sass:
        write_sass(a.tag)
    reports = a.reports if a.reports is not None else sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{a.tag}_*__*.csv")) +
                                                            glob.glob(os.path.join(ROOT, "gpurun_out", f"{a.tag}_*.ncu-rep")))
    traffic = {}
    for rep in reports:
        summarise(rep, a.tag, traffic)
    if traffic:
        for wl in traffic.values():
            for mode in wl.values():
                for k, rec in mode.items():
                    n = rec.pop("launches")
                    b, ns = rec.pop("dram_bytes"), rec.pop("ns")
                    rec.update(launches_captured=n, dram_bytes_per_launch=b / n, ncu_duration_us=ns / n / 1e3,
                               achieved_dram_gbs_under_ncu=b / ns if ns else None)
        path = os.path.join(ROOT, "profiles", f"{a.tag}_traffic.json")
        with open(path, "w") as f:
            json.dump(traffic, f, indent=1, sort_keys=True)
        print("wrote", path)


if __name__ == "__main__":
    main()
