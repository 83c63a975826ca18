python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "loss_and_gradient" 2>&1 | grep -E "^E|PER-TENSOR|passed|failed" | head -20
python -m pytest tests/test_gpu_path.py -m gpu -q -k "stream_update or early_fc" 2>&1 | tail -3
