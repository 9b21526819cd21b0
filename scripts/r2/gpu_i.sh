# round 2, run I (2 GPUs): splitter lookup table — multi-GPU parity (both modes), in-Thrill 2 workers, N=2 bench
set -x
export TG_DEBUG_EXCHANGE=1
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_nodes.py tests/test_gpu_sort_kernels.py -m gpu -q -x 2>&1 | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/r2i_bench_sort_n2.json 2> gpurun_out/r2i_bench_sort_n2.err; tail -3 gpurun_out/r2i_bench_sort_n2.err | cut -c1-500; cut -c1-2600 gpurun_out/r2i_bench_sort_n2.json
