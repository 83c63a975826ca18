python -m pytest tests/test_gpu_path.py tests/test_host_emulator.py tests/test_gpu_async.py -x -q 2>&1 | tail -5
for lean in 1 0; do
  echo "=== ARL_FRAME_LEAN=$lean"
  ARL_FRAME_LEAN=$lean python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'])"
  ARL_FRAME_LEAN=$lean python bench.py --workload frame_sweep 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['sweep']); print(d['rgb_sweep'])"
done
