timeout 900 python -m pytest tests/test_gpu_async.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2l_tests2.log
cat gpurun_out/r2l_tests2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --parallelism async --envs 64 --game space_invaders --poll-horizon 32 > gpurun_out/r2l_bench_async2.json 2> gpurun_out/r2l_bench_async2.err
tail -3 gpurun_out/r2l_bench_async2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2l_bench_async2.json').read().strip().splitlines()[-1])
print('async2', d['value'], d['ms_per_step'], d['phases'], d.get('parity'))
"
