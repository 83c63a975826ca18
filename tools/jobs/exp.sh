for v in "3 0" "0 0" "3 1" "3 0" "0 0" "3 1"; do
  set -- $v
  echo "=== ARL_L2_PERSIST=$1 ARL_OBS_L2_HINT=$2"
  ARL_L2_PERSIST=$1 ARL_OBS_L2_HINT=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'], {k['kernel']: round(k['ms']*1e3,2) for k in d['kernels'] if k['kernel'] in ('conv0_fwd','conv0_wgrad','clip_update')})"
done
