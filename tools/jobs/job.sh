python tools/sanitize_workload.py 2>&1 | tail -2
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== $tool"
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_workload.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_WORKLOAD_OK|hazard|Invalid|Uninitialized" gpurun_out/r2_sanitize_$tool.log | sort | uniq -c | head -12
done
