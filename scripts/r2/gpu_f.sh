# round 2, run F (2 GPUs): quick checks, whole GPU suite (multi-GPU parity both exchange modes, in-Thrill 2 workers, device
# Files), bench at N=1 and N=2 incl. the TeraSort extra, launch list of the reduce path
set -x
export TG_DEBUG_EXCHANGE=1
TG_DEBUG_REDUCE=1 timeout 90 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "quick_reduce zipf failed"; exit 1; }
timeout 90 python scripts/quick_sort.py 100000000 6
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "gpu tests failed"; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2f_bench_sort_n1.json 2> gpurun_out/r2f_bench_sort_n1.err; tail -3 gpurun_out/r2f_bench_sort_n1.err; cut -c1-9000 gpurun_out/r2f_bench_sort_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_sort_n2.json 2> gpurun_out/r2f_bench_sort_n2.err; tail -3 gpurun_out/r2f_bench_sort_n2.err | cut -c1-500; cut -c1-9000 gpurun_out/r2f_bench_sort_n2.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_reduce.csv python scripts/quick_reduce.py 125000000 3 > gpurun_out/r2f_reduce_under_ncu.log 2>&1
ls -la gpurun_out | tail
