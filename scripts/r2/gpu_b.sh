# round 2, run B (2 GPUs): GPU test suite incl. the multi-GPU parity worker (P2P exchange), NCCL-mode exchange, timings
set -x
export TG_DEBUG_EXCHANGE=1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40
TG_EXCHANGE=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/multi_gpu_worker.py 2>&1 | tail -5
timeout 300 python scripts/quick_sort.py 100000000 6
timeout 300 python scripts/quick_reduce.py 125000000 5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -5 gpurun_out/r2b_bench_n2.err; cut -c1-2500 gpurun_out/r2b_bench_n2.json
