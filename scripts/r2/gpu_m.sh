set -x
timeout 200 python -m pytest tests/test_gpu_devfile.py tests/test_gpu_host_nodes.py -m gpu -q -x 2>&1 | tail -8
