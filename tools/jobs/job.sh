python -m pytest tests/test_gpu_path.py -x -q -k "example_script" 2>&1 | tail -15
