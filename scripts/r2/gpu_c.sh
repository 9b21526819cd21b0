# round 2, run C (8 GPUs): multi-GPU parity worker (P2P exchange), bench lines at N=8 (tight timeouts: 8x charge)
set -x
export TG_DEBUG_EXCHANGE=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 tests/multi_gpu_worker.py > gpurun_out/r2c_parity_w8.log 2>&1; tail -4 gpurun_out/r2c_parity_w8.log | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err; tail -3 gpurun_out/r2c_bench_n8.err | cut -c1-600; cut -c1-6000 gpurun_out/r2c_bench_n8.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 --metric reduce --no-extras > gpurun_out/r2c_bench_reduce_n8.json 2> gpurun_out/r2c_bench_reduce_n8.err; tail -3 gpurun_out/r2c_bench_reduce_n8.err | cut -c1-600; cut -c1-4000 gpurun_out/r2c_bench_reduce_n8.json
