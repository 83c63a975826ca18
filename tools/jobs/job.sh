python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
ARL_FWD_WIDE=15 ARL_LIB_PATH=/root/repo/accel_rl_b200/csrc/libaccelrl_b200_trace.so python tests/trace_persist.py 2>&1 | tail -17 | head -10
for v in 15 0 6; do
  echo "=== ARL_FWD_WIDE=$v"
  ARL_FWD_WIDE=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases']); print({k['kernel']: round(k['ms']*1e3,2) for k in d['kernels'] if 'conv' in k['kernel']})"
done
