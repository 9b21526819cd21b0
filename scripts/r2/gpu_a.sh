# round 2, run A: baseline timings + ncu captures of the SHIPPED launch configuration (cfg2 = 256 threads x 3 CTAs/SM)
set -x
timeout 300 python scripts/quick_sort.py 100000000 6
timeout 300 python scripts/quick_reduce.py 125000000 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 8 -c 3 -o gpurun_out/r2a_partition -f python scripts/quick_sort.py 100000000 4 > gpurun_out/r2a_partition.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:prefix_fixup|chunk_hist|seg_count' -s 6 -c 3 -o gpurun_out/r2a_aux -f python scripts/quick_sort.py 100000000 4 > gpurun_out/r2a_aux.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_units -s 2 -c 1 -o gpurun_out/r2a_agg -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2a_agg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 4 -c 2 -o gpurun_out/r2a_rpart -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/r2a_rpart.log 2>&1
nvidia-smi topo -m | head -12
ls -la gpurun_out | tail -12
