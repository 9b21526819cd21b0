timeout 40 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x -k "very_frequent" 2>&1 | tail -3
