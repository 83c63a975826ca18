python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log
