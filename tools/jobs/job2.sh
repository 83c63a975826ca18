timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "2" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/timeline.py --sync 2>&1 | grep " us " | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sync2', d['value'], d['ms_per_step'], d['phases'], d.get('sync_timeline'))
"
