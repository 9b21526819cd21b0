# round 2, run P (2 GPUs): last multi-GPU validation (exchange plan refactor, reduce tuning): whole GPU suite + N=2 bench lines
set -x
export TG_DEBUG_EXCHANGE=1
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 --metric reduce --no-extras > gpurun_out/r2p_bench_reduce_n2.json 2> gpurun_out/r2p_bench_reduce_n2.err; tail -2 gpurun_out/r2p_bench_reduce_n2.err | cut -c1-300; cut -c1-400 gpurun_out/r2p_bench_reduce_n2.json
