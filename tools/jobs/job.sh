python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
