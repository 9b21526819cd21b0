timeout 200 python -m pytest tests/test_gpu_reduce.py -x -q 2>&1 | tail -4
TG_DEBUG_REDUCE=1 timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
