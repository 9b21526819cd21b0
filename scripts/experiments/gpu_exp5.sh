# segmented prefix sort: correctness + timing
timeout 600 python -m pytest tests/test_gpu_radix.py tests/test_gpu_sort_kernels.py -x -q 2>&1 | tail -8
for seg in 0 1; do for cfg in 0 2 4; do echo -n "seg=$seg "; TG_SEGMENTED=$seg TG_SWEEP_CFG=$cfg timeout 120 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1; done; done
