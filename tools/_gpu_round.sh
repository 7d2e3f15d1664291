timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r1_bench_1gpu_final.json 2> gpurun_out/bench_final.err; echo rc=$?
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2>> gpurun_out/bench_final.err; echo rc=$?
head -c 300 gpurun_out/r1_bench_reference_arm.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1_bench_1gpu_final.json'))
print(round(d['value']), round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['clocks'], d['gpu_launches'])
print(d['kernel_breakdown_ms_per_micro_batch'])
PY
