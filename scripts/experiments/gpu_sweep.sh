# partition-kernel launch-configuration sweep (timing experiment)
for cfg in 0 1 2 3 4 5 6 7; do TG_SWEEP_CFG=$cfg timeout 120 python scripts/quick_sort.py 100000000 6 2>&1 | tail -2; done
timeout 600 python -m pytest tests/test_gpu_radix.py tests/test_gpu_sort_kernels.py -x -q 2>&1 | tail -5
