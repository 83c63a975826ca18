python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "loss_and_gradient or nature" 2>&1 | grep -E "^E|PER-TENSOR|passed|failed" | head -20
for w in 1 0; do
ARL_WGRAD_WIDE=$w python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wide=$w', d['value'], d['ms_per_step'], d['phases'])
for k in d['kernels'][:12]: print('   ',k['kernel'],k['ms'],k['share'],k['tflops'])
"
done
