# round 2, run L (1 GPU): final state — GPU suite, smoke, both bench lines with the CPU baselines and the in-Thrill extra
set -x
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/r2l_bench_sort_n1.json 2> gpurun_out/r2l_bench_sort_n1.err; tail -3 gpurun_out/r2l_bench_sort_n1.err | cut -c1-400
timeout 400 python bench.py --metric reduce > gpurun_out/r2l_bench_reduce_n1.json 2> gpurun_out/r2l_bench_reduce_n1.err; tail -3 gpurun_out/r2l_bench_reduce_n1.err | cut -c1-400
python - <<'P'
import json
for f in ['gpurun_out/r2l_bench_sort_n1.json','gpurun_out/r2l_bench_reduce_n1.json']:
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, l['value'], l['ms_per_step'], 'frac', l['roofline']['frac'], json.dumps(l['roofline']['step_share']))
    print('  e2e', json.dumps(l['e2e'])[:260]); print('  cpu', json.dumps(l['cpu_baseline'])[:200])
    for k,v in l['extra'].items(): print('  extra', k, json.dumps(v)[:1200])
P
