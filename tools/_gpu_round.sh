TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR tools/vocab_parallel_check.py 2>&1 | grep -v "^W\|^\*\*\*\*" | tail -3
timeout 200 $TR tools/vocab_parallel_check.py --perf 2>&1 | grep -v "^W\|^\*\*\*\*" | tail -3
PRL_SAVE_LOGITS=0 timeout 200 $TR tools/vocab_parallel_check.py --perf 2>&1 | grep -v "^W\|^\*\*\*\*" | tail -1
timeout 200 $TR tools/fsdp_check.py 2>&1 | grep -v "^W\|^\*\*\*\*" | tail -3
timeout 300 $TR bench.py --gpus 2 --micro-batches 8 --no-cpu-baseline --no-breakdown --no-torch-baseline > gpurun_out/r1_bench_2gpu_8mb.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
head -c 600 gpurun_out/r1_bench_2gpu_8mb.json
