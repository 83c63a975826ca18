python -m pytest tests/test_gpu_path.py -x -q -k "reconfigure" 2>&1 | tail -5
