# where does the partition pass spend its time?  (DBG instantiation; results of dbg runs are wrong by design)
export TG_PREFIX_SORT=0
for dbg in 0 1 2 4 8 9 11 15; do echo "dbg=$dbg"; TG_SWEEP_DEBUG=$dbg timeout 120 python scripts/quick_sort.py 100000000 4 2>&1 | tail -1; done
unset TG_PREFIX_SORT
timeout 120 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 4 -c 1 -o gpurun_out/prof_sweep_r1d -f python scripts/quick_sort.py 100000000 3 > gpurun_out/prof_sweep_r1d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefix_fixup -s 1 -c 1 -o gpurun_out/prof_fixup_r1d -f python scripts/quick_sort.py 100000000 3 > gpurun_out/prof_fixup_r1d.log 2>&1
timeout 600 python -m pytest tests/test_gpu_radix.py tests/test_gpu_sort_kernels.py -x -q 2>&1 | tail -5
bash scripts/gpu_prof_reduce.sh
