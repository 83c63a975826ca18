python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -k "loss_and_gradient or a2c_nonreset or stream_update or early_fc" 2>&1 | tail -5 > gpurun_out/r2e_tests.log
for e in 0 1; do
ARL_EARLY_FC=$e python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_early$e.json 2> gpurun_out/r2e_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench_early$e.json').read().strip().splitlines()[-1])
print('early=$e', d['value'], d['ms_per_step'], d['phases'])
for k in d['kernels'][:6]: print('   ',k['kernel'],k['ms'],k['share'],k['tflops'])
"
done
tail -5 gpurun_out/r2e_tests.log
