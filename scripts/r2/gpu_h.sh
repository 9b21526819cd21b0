# round 2, run H (1 GPU): sync-free prefix sort, PageRank section, both bench lines, ncu summaries for profiles/
set -x
timeout 90 python scripts/quick_sort.py 100000000 6
TG_SORT_OPTIMISTIC=0 timeout 90 python scripts/quick_sort.py 100000000 6
TG_DEBUG_REDUCE=1 timeout 90 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/r2h_bench_sort_n1.json 2> gpurun_out/r2h_bench_sort_n1.err; tail -3 gpurun_out/r2h_bench_sort_n1.err; cut -c1-2500 gpurun_out/r2h_bench_sort_n1.json
N="--set full --clock-control none --import-source on"
T=/tmp/ncu; mkdir -p $T
timeout 200 ncu $N -k regex:partition_kernel -s 8 -c 3 -o $T/partition_u64 -f python scripts/quick_sort.py 100000000 4 > gpurun_out/r2h_ncu1.log 2>&1
timeout 200 ncu $N -k regex:partition_kernel -s 4 -c 2 -o $T/partition_kv16 -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2h_ncu2.log 2>&1
timeout 200 ncu $N -k 'regex:agg_units|hot_hist' -s 4 -c 2 -o $T/agg -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2h_ncu3.log 2>&1
timeout 300 ncu $N -k 'regex:partition_kernel|merge2_kernel' -c 12 -o $T/aux -f python scripts/r2/aux_kernels.py 40000000 > gpurun_out/r2h_ncu4.log 2>&1
for f in partition_u64 partition_kv16 agg aux; do
  timeout 120 python profiles/summarize.py kernel $T/$f.ncu-rep > gpurun_out/r2h_${f}_kernel.txt 2>&1
  timeout 120 python profiles/summarize.py source $T/$f.ncu-rep 40 > gpurun_out/r2h_${f}_source.txt 2>&1
done
cuobjdump -sass thrill_b200/csrc/libthrill_gpu.so 2>/dev/null | grep -E "UBLKCP|SYNCS|ATOMS|VOTE|REDUX|LDG.E.128|STG.E.128|MATCH" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c > gpurun_out/r2h_sass_mnemonics.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -5
