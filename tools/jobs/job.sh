for v in 148 100 116 132 148; do
  echo "=== ARL_WGRAD0_CTAS=$v"
  ARL_WGRAD0_CTAS=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'])"
done
