set -x
timeout 200 python bench.py --metric reduce --no-cpu-baseline > gpurun_out/r2r_bench_reduce_n1.json 2> gpurun_out/r2r_bench_reduce_n1.err; tail -2 gpurun_out/r2r_bench_reduce_n1.err; cut -c1-200 gpurun_out/r2r_bench_reduce_n1.json
