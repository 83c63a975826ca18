python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_gpu_tests.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 400 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_ref.json
python bench.py --algo a2c --envs 1024 --horizon 5 --game mix4 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c3_a2c_mix4.json 2> gpurun_out/r2_bench_c3.err; tail -c 300 gpurun_out/r2_bench_c3_a2c_mix4.json
python bench.py --workload frame_sweep --game mix4 > gpurun_out/r2_frame_sweep_mix4.json 2> gpurun_out/r2_fs.err; cat gpurun_out/r2_frame_sweep_mix4.json
python bench.py --frames rgb --no-cpu-baseline > gpurun_out/r2_bench_rgb.json 2> gpurun_out/r2_bench_rgb.err; tail -c 300 gpurun_out/r2_bench_rgb.json
