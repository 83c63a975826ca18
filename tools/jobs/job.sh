for v in 2 0 2 0; do
  echo "=== ARL_L2_PERSIST=$v"
  ARL_L2_PERSIST=$v python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline 2> gpurun_out/l2.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases']); print({k['kernel']: round(k['ms']*1e3,2) for k in d['kernels'][:8]})"
  grep "L2 persistence" gpurun_out/l2.err | head -1
done
python -m pytest tests/test_gpu_path.py -x -q -k "bit_identical or ppo_iteration" 2>&1 | tail -2
python bench.py --algo a2c --envs 1024 --horizon 5 --game mix4 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c3_a2c_mix4.json 2> gpurun_out/r2_bench_c3.err; tail -c 300 gpurun_out/r2_bench_c3_a2c_mix4.json
