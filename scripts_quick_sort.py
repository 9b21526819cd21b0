# scratch: quick device-resident timing of tg_radix_sort_local (not the bench contract)
import ctypes as C, sys
from thrill_b200 import capi
c = capi.Ctx(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
d = c.alloc(n * 8); tmp = c.alloc(n * 8)
desc = capi.u64_desc()
for i in range(6):
    c.ck(c.L.tg_gen_sort_uniform(c.h, d, 0, n, 42)); c.sync()
    c.timer_start()
    c.ck(c.L.tg_radix_sort_local(c.h, C.byref(desc), d, tmp, n))
    ms = c.timer_stop()
    print("iter", i, "ms", ms, "Gkeys/s", n / ms / 1e6, "GB/s(136B/key)", 136 * n / ms / 1e6, flush=True)
print("sorted", c.is_sorted(desc, d, n))
