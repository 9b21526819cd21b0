# scratch: runs the stand-alone kernel-level entry points at scale so that ncu can capture them (classify/scatter by
# splitters, hash partition by Hash128to64 % p, 2-way merge levels of tg_kway_merge)
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from thrill_b200 import capi
c = capi.Ctx(0)
L = c.L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
p = 8
d = c.alloc(n * 8 + 64); out = c.alloc(n * 8 + 64); tmp = c.alloc(n * 8 + 64)
desc = capi.u64_desc()
c.ck(L.tg_gen_sort_uniform(c.h, d, 0, n, 42)); c.sync()
# splitters: (key, index) pairs at the p-quantiles of the key space
spl = np.zeros((p - 1, 2), dtype=np.uint64)
for i in range(1, p):
    spl[i - 1, 0] = (i << 61)
    spl[i - 1, 1] = 12345
counts = (C.c_uint64 * p)()
for it in range(3):
    c.timer_start()
    c.ck(L.tg_classify_scatter(c.h, C.byref(desc), d, n, 0, spl.ctypes.data, p, out, counts))
    print("classify_scatter p=%d: %.3f ms  counts=%s" % (p, c.timer_stop(), [int(x) for x in counts][:4]), flush=True)
# k-way merge of p sorted runs (the received runs of the merge pipeline)
runs = (C.c_uint64 * p)(*[n // p] * p)
nn = (n // p) * p
off = 0
for r in range(p):
    c.ck(L.tg_radix_sort_local(c.h, C.byref(desc), d + off * 8, tmp, n // p))
    off += n // p
c.sync()
for it in range(2):
    c.timer_start()
    c.ck(L.tg_kway_merge(c.h, C.byref(desc), d, runs, p, out, tmp))
    print("kway_merge k=%d of %d keys: %.3f ms" % (p, nn, c.timer_stop()), flush=True)
assert c.is_sorted(desc, out, nn)
c.free(d); c.free(out); c.free(tmp)
# hash partition of 16-byte records
m = 125000000
dk = c.alloc(m * 16); ok = c.alloc(m * 16)
c.ck(L.tg_gen_reduce_uniform(c.h, dk, 0, m, 42, 1 << 26, 0)); c.sync()
kvd = capi.KVDesc(16, capi.OP_SUM_F64)
for it in range(3):
    c.timer_start()
    c.ck(L.tg_hash_partition(c.h, C.byref(kvd), dk, m, p, ok, counts))
    print("hash_partition p=%d: %.3f ms" % (p, c.timer_stop()), flush=True)
