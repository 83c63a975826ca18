python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 200 gpurun_out/r2_bench_ref.json
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 48 -f -o /tmp/prof_r2_full python tools/ncu_workload.py --horizon 2 > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log
python tools/summarize_ncu.py full /tmp/prof_r2_full.ncu-rep gpurun_out/r2_full.md 2>&1 | tail -2
ncu -i /tmp/prof_r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ls -la /tmp/prof_r2_full.ncu-rep
sz=$(stat -c %s /tmp/prof_r2_full.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_r2_full.ncu-rep gpurun_out/; fi
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; tail -2 gpurun_out/r2_ncu_launches.log; wc -l gpurun_out/r2_launches.csv
du -sh gpurun_out
