timeout 280 python bench.py > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; tail -3 gpurun_out/bench_n1_full.err; cut -c1-400 gpurun_out/bench_n1_full.json
