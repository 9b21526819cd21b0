# ncu captures of the ReduceByKey kernels + plain timings
timeout 300 python scripts/quick_reduce.py 125000000 5
timeout 300 python scripts/quick_reduce.py 125000000 5 uniform
for k in preagg_kernel aggregate_kernel compact_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_${k}_r1d -f python scripts/quick_reduce.py 125000000 4 > gpurun_out/prof_${k}.log 2>&1
done
ls -la gpurun_out
