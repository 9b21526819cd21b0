set -x
timeout 60 python scripts/quick_sort.py 100000000 6
TG_UNSTABLE_CFG=9 timeout 60 python scripts/quick_sort.py 100000000 6
TG_UNSTABLE_CFG=8 timeout 60 python scripts/quick_sort.py 100000000 6
TG_SWEEP_CFG=9 timeout 60 python scripts/quick_sort.py 100000000 6
