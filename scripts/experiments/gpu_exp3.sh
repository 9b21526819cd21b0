# look-back batch size x launch configuration (8 plain LSD passes, prefix sort off)
export TG_PREFIX_SORT=0
for v in lb8pf1 lb16pf0 lb32pf0 lb32pf1 lb64pf0; do
  for cfg in 0 1 2 3 4 6; do
    echo -n "$v "; TG_LIB=$PWD/variants/$v/libthrill_gpu.so TG_SWEEP_CFG=$cfg timeout 120 python scripts/quick_sort.py 100000000 4 2>&1 | tail -1
  done
done
