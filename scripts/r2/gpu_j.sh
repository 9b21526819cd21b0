# round 2, run J (2 GPUs): record exchange with 128-bit stores — multi-GPU parity (TeraSort 1e6 golden), N=2 bench with extras
set -x
export TG_DEBUG_EXCHANGE=1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_sort_n2.json 2> gpurun_out/r2j_bench_sort_n2.err; tail -3 gpurun_out/r2j_bench_sort_n2.err | cut -c1-500; python - <<'P'
import json
l=json.loads(open('gpurun_out/r2j_bench_sort_n2.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], json.dumps(l['roofline']['step_share']))
print(json.dumps(l['extra'].get('terasort'))[:1500])
print(json.dumps(l['extra'].get('reduce'))[:700])
P
