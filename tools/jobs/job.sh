python -m pytest tests/test_gpu_path.py -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  echo "=== ARL_FC_CLUSTER=$v"
  ARL_FC_CLUSTER=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2>gpurun_out/fcc.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases']); print({k['kernel']: round(k['ms']*1e3,2) for k in d['kernels'] if 'fc_fwd' in k['kernel'] or 'head' in k['kernel']})" || grep -v "^frame" gpurun_out/fcc.err | tail -5
done
