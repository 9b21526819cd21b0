set -x
TG_DEBUG_REDUCE=1 timeout 60 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
timeout 60 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --metric reduce --no-cpu-baseline --no-extras > gpurun_out/r2o_bench_reduce_n1.json 2> gpurun_out/r2o_bench_reduce_n1.err; tail -2 gpurun_out/r2o_bench_reduce_n1.err; cut -c1-300 gpurun_out/r2o_bench_reduce_n1.json
