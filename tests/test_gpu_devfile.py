"""Device-resident Files and the small bookkeeping entry points of the C ABI, on one GPU: an operator result
detached as a tg_dev_file feeds the next operator without crossing PCIe; the transfer counters prove it."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from gpu_util import make_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from thrill_b200 import capi
    c = capi.Ctx(0)
    yield c
    c.close()


def _transfer(ctx):
    h, d = C.c_uint64(), C.c_uint64()
    ctx.ck(ctx.L.tg_transfer_bytes(ctx.h, C.byref(h), C.byref(d)))
    return h.value, d.value


def test_sort_then_reduce_through_a_device_file(ctx):
    from thrill_b200 import capi
    n = 600000
    kv = O.gen_reduce_zipf(0, n, O.zipf_cdf(5000), exact=1)
    blocks, nb, raw = make_blocks(capi, kv, 1 << 18)
    h0, d0 = _transfer(ctx)
    n_sorted = C.c_size_t()
    ctx.ck(ctx.L.tg_sort_file(ctx.h, C.byref(capi.kv_key_desc()), blocks, nb, 3, C.byref(n_sorted)))
    assert n_sorted.value == n
    f = capi.DevFile()
    ctx.ck(ctx.L.tg_output_detach(ctx.h, C.byref(f)))
    assert f.items == n and f.item_bytes == 16 and f.dptr
    h1, d1 = _transfer(ctx)
    assert (h1 - h0, d1 - d0) == (n * 16, 0)                    # the input went up once, nothing came down
    # the device File is an ordinary sorted File: fetch a copy (lazy D2H) and compare with the oracle's stable sort
    out = np.zeros(n, dtype=O.KV)
    ob, nob, _ = make_blocks(capi, out, 1 << 20)
    ctx.ck(ctx.L.tg_dev_file_fetch(ctx.h, C.byref(f), ob, nob))
    assert np.array_equal(out, np.ascontiguousarray(O.sort_items(kv, O.KV_DESC)).view(O.KV).reshape(-1))
    h2, d2 = _transfer(ctx)
    assert (h2 - h1, d2 - d1) == (0, n * 16)
    # ... and the input of the next operator, twice (the handle stays intact: a DIA may have several children)
    for _ in range(2):
        n_red = C.c_size_t()
        ctx.ck(ctx.L.tg_reduce_dev(ctx.h, C.byref(capi.KVDesc(16, capi.OP_SUM_F64)), C.byref(f), C.byref(n_red)))
        red = np.zeros(n_red.value, dtype=O.KV)
        rb, nrb, _ = make_blocks(capi, red, 1 << 20)
        ctx.ck(ctx.L.tg_fetch_output(ctx.h, rb, nrb))
        assert np.array_equal(np.sort(red, order="key"), O.reduce_simple(kv, O.OP_SUM_F64))
    h3, d3 = _transfer(ctx)
    assert h3 - h2 == 0 and d3 - d2 == 2 * n_red.value * 16     # no upload between the operators
    # sort of a device File
    n_s2 = C.c_size_t()
    ctx.ck(ctx.L.tg_sort_dev(ctx.h, C.byref(capi.kv_key_desc()), C.byref(f), 9, C.byref(n_s2)))
    out2 = np.zeros(n, dtype=O.KV)
    ob2, nob2, _ = make_blocks(capi, out2, 1 << 20)
    ctx.ck(ctx.L.tg_fetch_output(ctx.h, ob2, nob2))
    assert np.array_equal(out2, out)
    ctx.ck(ctx.L.tg_dev_file_free(ctx.h, C.byref(f)))
    assert not f.dptr
    # wrong item size is refused
    g = capi.DevFile(None, 0, 8, 0)
    st = ctx.L.tg_reduce_dev(ctx.h, C.byref(capi.KVDesc(16, capi.OP_SUM_F64)), C.byref(g), C.byref(n_red))
    assert st != 0


def test_profile_list_and_device_count(ctx):
    from thrill_b200 import capi
    assert ctx.L.tg_device_count() >= 1
    n = 2000000
    d = ctx.alloc(n * 8); tmp = ctx.alloc(n * 8)
    ctx.ck(ctx.L.tg_gen_sort_uniform(ctx.h, d, 0, n, 1))
    ctx.profile_enable(True)
    ctx.ck(ctx.L.tg_radix_sort_local(ctx.h, C.byref(capi.u64_desc()), d, tmp, n))
    tot, cnt = ctx.profile_get(capi.K_PARTITION)
    lst = ctx.profile_list(capi.K_PARTITION)
    ctx.profile_enable(False)
    assert cnt == len(lst) and cnt >= 3 and abs(sum(lst) - tot) < 1e-3 * max(tot, 1e-3) + 1e-4
    assert ctx.is_sorted(capi.u64_desc(), d, n)
    ctx.free(d); ctx.free(tmp)
