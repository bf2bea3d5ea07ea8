timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -x -q -s --durations=5 2>&1 | grep -v "^$" | tail -30
