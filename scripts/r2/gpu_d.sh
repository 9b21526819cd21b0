# round 2, run D (1 GPU): GPU tests, both bench lines, ncu captures of every kernel class in its shipped configuration
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/r2d_bench_sort_n1.json 2> gpurun_out/r2d_bench_sort_n1.err; tail -3 gpurun_out/r2d_bench_sort_n1.err; cut -c1-5000 gpurun_out/r2d_bench_sort_n1.json
timeout 600 python bench.py --metric reduce > gpurun_out/r2d_bench_reduce_n1.json 2> gpurun_out/r2d_bench_reduce_n1.err; tail -3 gpurun_out/r2d_bench_reduce_n1.err; cut -c1-4000 gpurun_out/r2d_bench_reduce_n1.json
timeout 300 python scripts/r2/aux_kernels.py
N="--set full --clock-control none --import-source on"
timeout 600 ncu $N -k regex:partition_kernel -s 8 -c 2 -o gpurun_out/r2d_partition_u64 -f python scripts/quick_sort.py 100000000 4 > gpurun_out/r2d_ncu1.log 2>&1
timeout 600 ncu $N -k regex:partition_kernel -s 4 -c 2 -o gpurun_out/r2d_partition_kv16 -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2d_ncu2.log 2>&1
timeout 600 ncu $N -k regex:agg_units -s 2 -c 1 -o gpurun_out/r2d_agg -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2d_ncu3.log 2>&1
timeout 600 ncu $N -k 'regex:partition_kernel|merge2_kernel' -c 40 -o gpurun_out/r2d_aux -f python scripts/r2/aux_kernels.py 20000000 > gpurun_out/r2d_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
