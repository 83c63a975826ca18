# 8-GPU tests + synchronous N = 8 / 4 and asynchronous 8-GPU bench lines (gpurun --gpus 8 --timeout 1500 -- 'mkdir -p gpurun_out; bash tools/jobs/job8.sh')
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "4 or 8" 2>&1 | tail -6 > gpurun_out/r2_tests8.log
cat gpurun_out/r2_tests8.log
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
tail -2 gpurun_out/r2_bench_n$n.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n$n.json').read().strip().splitlines()[-1])
print('sync$n', d['value'], d['ms_per_step'], d['phases'], d.get('parity'), d.get('sync_timeline'), d.get('e2e'), d.get('roofline'))
"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 4 --no-e2e --no-cpu-baseline --parallelism async --envs 64 --game space_invaders --poll-horizon 32 > gpurun_out/r2_bench_async8.json 2> gpurun_out/r2_bench_async8.err
tail -2 gpurun_out/r2_bench_async8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_async8.json').read().strip().splitlines()[-1])
print('async8', d['value'], d['ms_per_step'], d['phases'], d.get('parity'))
"
