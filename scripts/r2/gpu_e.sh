# round 2, run E (1 GPU): quick checks first (tight timeouts: a hang must cost seconds, not minutes), then tests, bench lines,
# ncu captures summarised on the box (the .ncu-rep files stay there: 64 MiB limit)
set -x
TG_DEBUG_REDUCE=1 timeout 90 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -4 || exit 1
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "quick_reduce zipf failed"; exit 1; }
timeout 90 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -2
timeout 90 python scripts/quick_sort.py 100000000 6
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "gpu tests failed"; exit 1; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench_sort_n1.json 2> gpurun_out/r2e_bench_sort_n1.err; tail -3 gpurun_out/r2e_bench_sort_n1.err; cut -c1-1500 gpurun_out/r2e_bench_sort_n1.json
timeout 300 python bench.py --metric reduce --no-cpu-baseline > gpurun_out/r2e_bench_reduce_n1.json 2> gpurun_out/r2e_bench_reduce_n1.err; tail -3 gpurun_out/r2e_bench_reduce_n1.err; cut -c1-3000 gpurun_out/r2e_bench_reduce_n1.json
N="--set full --clock-control none --import-source on"
T=/tmp/ncu; mkdir -p $T
timeout 200 ncu $N -k regex:partition_kernel -s 8 -c 3 -o $T/partition_u64 -f python scripts/quick_sort.py 100000000 4 > gpurun_out/r2e_ncu1.log 2>&1
timeout 200 ncu $N -k regex:partition_kernel -s 4 -c 2 -o $T/partition_kv16 -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2e_ncu2.log 2>&1
timeout 200 ncu $N -k 'regex:agg_units|hot_hist' -s 4 -c 2 -o $T/agg -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2e_ncu3.log 2>&1
for f in partition_u64 partition_kv16 agg; do
  timeout 120 python profiles/summarize.py kernel $T/$f.ncu-rep > gpurun_out/r2e_${f}_kernel.txt 2>&1
  timeout 120 python profiles/summarize.py source $T/$f.ncu-rep 50 > gpurun_out/r2e_${f}_source.txt 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
