# scratch: quick device-resident timing of tg_radix_sort_local with the per-kernel-class profile (not the bench contract)
import ctypes as C, sys
from thrill_b200 import capi
c = capi.Ctx(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
d = c.alloc(n * 8 + (4 << 20)); tmp = c.alloc(n * 8 + (4 << 20))
desc = capi.u64_desc()
for i in range(6):
    c.ck(c.L.tg_gen_sort_uniform(c.h, d, 0, n, 42)); c.sync()
    if i == 3: c.profile_enable(True)
    c.timer_start()
    c.ck(c.L.tg_radix_sort_local(c.h, C.byref(desc), d, tmp, n))
    ms = c.timer_stop()
    print("iter", i, "ms %.3f" % ms, "Gkeys/s %.2f" % (n / ms / 1e6), flush=True)
pm, pc = c.profile_get(capi.K_PARTITION); hm, hc = c.profile_get(capi.K_RADIX_HIST)
print("partition ms/launch %.4f (%d) -> %.0f GB/s ; hist ms %.4f" % (pm / max(pc, 1), pc, 16 * n / (pm / max(pc, 1)) / 1e6, hm / max(hc, 1)))
print("sorted", c.is_sorted(desc, d, n))
