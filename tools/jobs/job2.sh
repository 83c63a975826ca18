# 2-GPU tests + synchronous bench line (gpurun --gpus 2 --timeout 1200 -- 'mkdir -p gpurun_out; bash tools/jobs/job2.sh')
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_async.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.err; tail -c 600 gpurun_out/r2_bench_n2.json
