timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "offline_evaluation" 2>&1 | tail -12
