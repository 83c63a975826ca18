python -m pytest tests/test_gpu_path.py -x -q -k "bit_identical or ppo_iteration or operand or runner" 2>&1 | tail -5
for v in 1 0 1 0; do
  echo "=== ARL_SPLIT_UPDATE=$v"
  ARL_SPLIT_UPDATE=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'])"
done
