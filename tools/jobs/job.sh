for d in 0 100 92; do
ARL_DGRAD_CTAS=$d python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dgrad_ctas=$d', d['value'], d['ms_per_step'], d['phases'])
"
done
ARL_DGRAD_CTAS=100 python tools/timeline.py 2>&1 | tail -18
