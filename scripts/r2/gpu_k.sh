# round 2, run K (4 GPUs): parity worker + bench sort line with extras (TeraSort exchange with staggered destinations)
set -x
export TG_DEBUG_EXCHANGE=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29608 tests/multi_gpu_worker.py > gpurun_out/r2k_parity_w4.log 2>&1; tail -2 gpurun_out/r2k_parity_w4.log | cut -c1-300
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2k_bench_sort_n4.json 2> gpurun_out/r2k_bench_sort_n4.err; tail -3 gpurun_out/r2k_bench_sort_n4.err | cut -c1-400
python - <<'P'
import json
l=json.loads(open('gpurun_out/r2k_bench_sort_n4.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], json.dumps(l['roofline']['step_share']))
print('e2e', json.dumps(l['e2e'])[:200])
for k,v in l['extra'].items(): print(k, json.dumps(v)[:900])
P
