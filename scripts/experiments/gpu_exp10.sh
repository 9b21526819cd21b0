timeout 200 python -m pytest tests/test_gpu_reduce.py -x -q 2>&1 | tail -4
timeout 150 python scripts/quick_reduce.py 125000000 5 2>&1 | tail -1
timeout 150 python scripts/quick_reduce.py 125000000 5 uniform 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 6 -c 1 -o gpurun_out/prof_segpass_r1e -f python scripts/quick_sort.py 100000000 3 > gpurun_out/prof_segpass_r1e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_units -s 2 -c 1 -o gpurun_out/prof_agg_units_r1f -f python scripts/quick_reduce.py 125000000 4 uniform > gpurun_out/prof_agg_units_r1f.log 2>&1
