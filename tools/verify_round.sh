#!/usr/bin/env bash
# What the round-end check runs on a B200 box (use through gpurun):  GPU parity tests, smoke(), the bench (both arms).
# Keep what is written under gpurun_out/ small: gpurun returns at most 64 MiB.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json; echo "reference arm rc=$?"
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "b200 arm rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["clocks"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
print({m: (r["value"], r["e2e"]["value"]) for m, r in d["modes"].items()})
PY
