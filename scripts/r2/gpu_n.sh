set -x
TG_DEBUG_REDUCE=1 timeout 60 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
TG_REDUCE_GROUP=0 timeout 60 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
TG_REDUCE_UNSTABLE=0 timeout 60 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
timeout 60 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_devfile.py -m gpu -q -x 2>&1 | tail -4
