set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for w in qwen3-1.7b-wordle qwen3-8b-deepscaler qwen3-0.6b-reverse; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r1_bench_1gpu_$w.json 2> gpurun_out/bench_$w.err; echo rc=$?
  python -c "
import json,sys
d=json.load(open('gpurun_out/r1_bench_1gpu_$w.json'))
print('$w', round(d['value']), round(d['e2e']['value']), d['frac_of_peak'], d['config']['lm_head_backward'][:20], d.get('kernel_breakdown_ms_per_micro_batch'))
"
done
