timeout 50 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_n1_extras.json 2> gpurun_out/bench_n1_extras.err; tail -2 gpurun_out/bench_n1_extras.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_extras.json')); print(d['value'], d['extra']['reduce_records_per_s'], d['extra'].get('reduce_uniform'))"
