# fast path (chunked MSD + segmented) + partitioned aggregation: correctness + timing
timeout 300 python -m pytest tests/test_gpu_radix.py tests/test_gpu_sort_kernels.py tests/test_gpu_reduce.py -x -q 2>&1 | tail -12
timeout 100 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1
TG_SWEEP_CFG=2 timeout 100 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1
timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -2
