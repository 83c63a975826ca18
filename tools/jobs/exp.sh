for v in 3 0 3 0 3 0; do
  echo "=== ARL_L2_PERSIST=$v"
  ARL_L2_PERSIST=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'], d['e2e']['value'])"
done
