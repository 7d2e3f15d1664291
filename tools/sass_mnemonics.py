"""Per-kernel SASS mnemonic counts of libprl_b200.so (cuobjdump -sass): the evidence that the hot kernels use
tcgen05 (UTCHMMA / LDTM / UTCBAR), TMA (UTMALDG / UTMASTG / UTMAREDG) and packed fp32x2 math."""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "prime_rl_cse_b200", "libprl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|FFMA2|FADD2|FMUL2|MUFU|LDG|STG|LDS|STS|RED|ATOM|BAR|SHFL|F2FP|UBLKCP|UCGABAR|ELECT)")
print("# SASS evidence for libprl_b200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,\n"
      "# UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor load / store / reduce-add (TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier,\n"
      "# UTCATOMSWS = tcgen05.alloc/dealloc, FFMA2 / FADD2 / FMUL2 = packed fp32x2 math.  gemm_kernel<CG, MS, A_MN, B_MN, EPI>;\n"
      "# EPI 0 = forward statistics, 1 = dlogits recompute, 2 = store / reduce-add, 3 = forward statistics + bf16 logits store.\n")
name, counts, n = None, None, 0


def flush():
    if name:
        print(name)
        print("   instructions=%d  " % n + "  ".join(f"{k}={v}" for k, v in sorted(counts.items())))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        n += 1
        op = m.group(1)
        if KEEP.match(op):
            counts[op] += 1
flush()
