# 8-GPU validation: collective operators vs the reference outputs, then the bench line at N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 tests/multi_gpu_worker.py 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err; cat gpurun_out/bench_n8.json | cut -c1-3000
