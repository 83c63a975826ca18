# 1-GPU regression + bench lines (run with: gpurun --timeout 1500 -- 'mkdir -p gpurun_out; bash tools/jobs/job.sh')
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 400 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_ref.json
