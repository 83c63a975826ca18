timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_async.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('sync2', d['value'], d['ms_per_step'], d['phases'], d.get('parity'), d.get('sync_timeline'), d.get('e2e'))
"
