# final single-GPU validation: GPU suite, bench line, ncu launch list of the bench command
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -2 gpurun_out/bench_n1_final.err; cut -c1-1500 gpurun_out/bench_n1_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
