python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "loss_and_gradient or nature or ppo_iteration or stream_update or operand_copies" 2>&1 | grep -E "^E|passed|failed" | head -20
python tools/timeline.py 2>&1 | grep " us " | tail -20
for w in 1 0; do
ARL_FUSED_BWD=$w python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fused_bwd=$w', d['value'], d['ms_per_step'], d['phases'])
for k in d['kernels'][:10]: print('   ',k['kernel'],k['ms'],k['share'],k['tflops'])
"
done
