set -x
python bench.py > gpurun_out/r1_bench_1gpu_saved.json 2> gpurun_out/bench_saved.err; echo rc=$?
CMD="python bench.py --no-cpu-baseline --no-breakdown --no-torch-baseline --micro-batches 8 --steps 1 --warmup 3"
$CMD > gpurun_out/bench_8mb.json 2> gpurun_out/bench_8mb.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|merge_partials|loss_|advantages_|dz_from_saved" -s 171 -c 57 --csv --log-file gpurun_out/r1_bench_launches_saved.csv $CMD > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo rc=$?
NCU_SAVED=1 NCU_V=151936 NCU_ITERS=1 python tools/ncu_case.py; echo rc=$?
NCU_SAVED=1 NCU_V=151936 NCU_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|dz_from_saved" -c 4 -o gpurun_out/r1_saved_logits -f python tools/ncu_case.py > gpurun_out/ncu_saved.log 2>&1; echo rc=$?
python tools/ncu_summary.py gpurun_out/r1_saved_logits.ncu-rep > gpurun_out/r1_saved_logits_ncu_summary.txt 2>&1
cat gpurun_out/r1_saved_logits_ncu_summary.txt
head -c 1500 gpurun_out/r1_bench_1gpu_saved.json
