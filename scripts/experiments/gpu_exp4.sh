# partition pass without the look-back wait (dbg bit0; wrong results by design) across launch configurations,
# plus the slimmed finishing pass
export TG_PREFIX_SORT=0
for cfg in 0 1 2 3 4 6; do
  for dbg in 1 3; do echo -n "dbg=$dbg "; TG_LIB=$PWD/variants/dbg/libthrill_gpu.so TG_SWEEP_DEBUG=$dbg TG_SWEEP_CFG=$cfg timeout 120 python scripts/quick_sort.py 100000000 4 2>&1 | tail -1; done
done
unset TG_PREFIX_SORT
timeout 120 python scripts/quick_sort.py 100000000 6 2>&1 | tail -1
