for v in 1 0 1 0; do
  echo "=== ARL_L2_PERSIST=$v"
  ARL_L2_PERSIST=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2> gpurun_out/l2.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'])"
done
