python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in 1 0; do
ARL_U8_CONV0=$w python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('u8_conv0=$w', d['value'], d['ms_per_step'], d['phases'])
for k in d['kernels'][:22]:
    if 'conv0' in k['kernel'] or 'frame' in k['kernel']: print('   ',k['kernel'],k['ms'],k['share'],k['tflops'])
"
done
