#!/usr/bin/env bash
# What the round-end check runs on a B200 box (use through gpurun):  GPU parity tests, smoke(), the bench (both arms).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json; echo "reference arm rc=$?"
timeout 400 python bench.py > gpurun_out/bench_1gpu.json; echo "b200 arm rc=$?"
head -c 400 gpurun_out/bench_1gpu.json; echo
