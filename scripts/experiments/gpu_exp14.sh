timeout 200 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_radix.py tests/test_gpu_sort_kernels.py -x -q 2>&1 | tail -3
timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
for cfg in 0 2 1; do TG_SWEEP_CFG=$cfg timeout 100 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1; done
