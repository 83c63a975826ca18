timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "4 or 8" 2>&1 | tail -15 > gpurun_out/r2m_tests8.log
cat gpurun_out/r2m_tests8.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2m_bench8.json 2> gpurun_out/r2m_bench8.err
tail -2 gpurun_out/r2m_bench8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench8.json').read().strip().splitlines()[-1])
print('sync8', d['value'], d['ms_per_step'], d['phases'], d.get('parity'), d.get('sync_timeline'))
"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --parallelism async --envs 64 --game space_invaders --poll-horizon 32 > gpurun_out/r2m_bench_async8.json 2> gpurun_out/r2m_bench_async8.err
tail -2 gpurun_out/r2m_bench_async8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench_async8.json').read().strip().splitlines()[-1])
print('async8', d['value'], d['ms_per_step'], d['phases'], d.get('parity'))
"
