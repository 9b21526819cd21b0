timeout 200 python -m pytest tests/test_gpu_reduce.py -x -q 2>&1 | tail -3
timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
TG_SWEEP_CFG=2 timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
TG_SWEEP_CFG=2 timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
TG_SWEEP_CFG=10 timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
