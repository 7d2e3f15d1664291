timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "k_split or saved or micro_step or engine_variants or backward_output or skips_masked or forward_active or vocab_sharded_backward or sampled_slices or hf_qwen3 or sft" 2>&1 | tail -6
for w in qwen3-0.6b-reverse qwen3-1.7b-wordle; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-torch-baseline > gpurun_out/r1_bench_1gpu_$w.json 2> gpurun_out/bench_$w.err; echo rc=$?
  python -c "
import json
d=json.load(open('gpurun_out/r1_bench_1gpu_$w.json'))
print('$w', round(d['value']), round(d['e2e']['value']), round(d['frac_of_peak']['measured_burst'],3), d.get('kernel_breakdown_ms_per_micro_batch'))
"
done
timeout 300 python bench.py --micro-batches 8 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_8mb_ks.json 2> gpurun_out/bench_8mb_ks.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_8mb_ks.json'))
print('headline-8mb', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], d.get('kernel_breakdown_ms_per_micro_batch'))
"
