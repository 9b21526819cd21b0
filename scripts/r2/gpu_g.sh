# round 2, run G (1 GPU): reduce changes (digit bytes, parallel unit list), GPU suite, both bench lines with the CPU baselines
set -x
TG_DEBUG_REDUCE=1 timeout 90 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "quick_reduce zipf failed"; exit 1; }
timeout 90 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
timeout 90 python scripts/quick_sort.py 100000000 6
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 400 python bench.py --metric reduce > gpurun_out/r2g_bench_reduce_n1.json 2> gpurun_out/r2g_bench_reduce_n1.err; tail -3 gpurun_out/r2g_bench_reduce_n1.err; cut -c1-3500 gpurun_out/r2g_bench_reduce_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_reduce.csv python scripts/quick_reduce.py 125000000 3 > gpurun_out/r2g_reduce_under_ncu.log 2>&1
