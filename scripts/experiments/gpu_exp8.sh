timeout 200 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_radix.py -x -q 2>&1 | tail -5
timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -2
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -2
for cfg in 0 2 8 9 10; do TG_SWEEP_CFG=$cfg timeout 100 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_units -s 2 -c 1 -o gpurun_out/prof_agg_units_r1e -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/prof_agg_units.log 2>&1
