timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -s -k "extreme or empty_and" 2>&1 | grep -v "^$" | tail -12
