timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2h_tests2.log
cat gpurun_out/r2h_tests2.log
for ov in 1 0; do
ARL_SYNC_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench2_ov$ov.json 2> gpurun_out/r2h_bench2_ov$ov.err
tail -2 gpurun_out/r2h_bench2_ov$ov.err
python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench2_ov$ov.json').read().strip().splitlines()[-1])
print('overlap=$ov', d['value'], d['ms_per_step'], d['phases'], d.get('parity'))
"
done
