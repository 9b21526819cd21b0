"""GPU parity: tg_radix_sort_local (hand-written sm_100a onesweep LSB radix sort) vs the oracle's
std::sort restatement (api/sort.hpp:789-796) — bit-exact.  Runs on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from thrill_b200 import capi
    c = capi.Ctx(device=0)
    yield c
    c.close()


def _sort_on_gpu(ctx, host, desc):
    from thrill_b200 import capi
    n = host.nbytes // desc.item_bytes
    d = ctx.to_device(host)
    tmp = ctx.alloc(max(host.nbytes, 16))
    ctx.ck(ctx.L.tg_radix_sort_local(ctx.h, C.byref(desc), d, tmp, n))
    out = ctx.download(d, host.nbytes)
    ctx.free(d); ctx.free(tmp)
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 8191, 8192, 8193, 100000, 1 << 20, 3000001])
def test_radix_u64_uniform_bit_exact(ctx, n):
    from thrill_b200 import capi
    keys = O.gen_sort_uniform(0, n) if n else np.empty(0, dtype=np.uint64)
    out = _sort_on_gpu(ctx, keys, capi.u64_desc()).view(np.uint64)
    assert np.array_equal(out, O.sort_items(keys).view(np.uint64))


@pytest.mark.parametrize("case", ["all_equal", "four_values", "reversed", "zipf", "high_bytes_only", "max_values"])
def test_radix_u64_edge_distributions(ctx, case):
    from thrill_b200 import capi
    n = 500000
    if case == "all_equal":
        keys = np.full(n, 1, dtype=np.uint64)                      # tests/api/sort_node_test.cpp:162-185
    elif case == "four_values":
        keys = (np.arange(n) % 4).astype(np.uint64)                # :187-210
    elif case == "reversed":
        keys = (n - np.arange(n) - 1).astype(np.uint64)            # :25-53
    elif case == "zipf":
        keys = O.gen_sort_zipf(0, n, O.zipf_cdf(1 << 16))
    elif case == "high_bytes_only":
        keys = (np.random.RandomState(1).randint(0, 1 << 16, size=n).astype(np.uint64)) << np.uint64(48)
    else:
        keys = np.full(n, 2**64 - 1, dtype=np.uint64); keys[::7] = 0
    out = _sort_on_gpu(ctx, keys, capi.u64_desc()).view(np.uint64)
    assert np.array_equal(out, np.sort(keys))


def test_radix_u64_descending(ctx):
    from thrill_b200 import capi
    keys = O.gen_sort_uniform(7, 200000)
    out = _sort_on_gpu(ctx, keys, capi.u64_desc(descending=True)).view(np.uint64)
    assert np.array_equal(out, np.sort(keys)[::-1])


@pytest.mark.parametrize("n", [5, 4097, 300000])
def test_radix_pairs_stable_by_key(ctx, n):
    """16-byte items sorted by their u64 key: equal keys keep input order (SortStable, api/sort.hpp:873-937),
    which is exactly what the oracle's stable merge sort produces."""
    from thrill_b200 import capi
    rng = np.random.RandomState(n)
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = rng.randint(0, 1000, size=n)
    kv["val"] = np.arange(n)
    out = _sort_on_gpu(ctx, kv, capi.kv_key_desc()).view(O.KV)
    ref = O.sort_items(kv, O.KV_DESC).view(O.KV)
    assert np.array_equal(out, ref)


def test_radix_tuple_big_endian_10_byte_key(ctx):
    """16-byte tuples {10-byte big-endian key, 2 pad, u32 index}: the TeraSort sort tuple
    (examples/terasort/terasort.cpp:35-37 comparator)."""
    from thrill_b200 import capi
    n = 100000
    rng = np.random.RandomState(4)
    t = np.zeros((n, 16), dtype=np.uint8)
    t[:, :10] = rng.randint(0, 256, size=(n, 10))
    t[:, 0] = rng.randint(0, 3, size=n)                 # many equal leading bytes
    t[:, 12:16] = np.arange(n, dtype=np.uint32).view(np.uint8).reshape(n, 4)
    desc = capi.KeyDesc(16, 0, 10, capi.KEY_BYTES_BE, 0, 0)
    out = _sort_on_gpu(ctx, t, desc).reshape(n, 16)
    ref = O.sort_items(t, O.KeyDesc(16, 0, 10, O.KEY_BYTES_BE)).reshape(n, 16)
    assert np.array_equal(out, ref)


def test_radix_full_size_properties(ctx):
    """cfg2 size (1e8 u64): sortedness + multiset preservation (order-independent checksum), generated and
    checked on the device."""
    from thrill_b200 import capi
    n = 100000000
    d = ctx.alloc(n * 8); tmp = ctx.alloc(n * 8)
    ctx.ck(ctx.L.tg_gen_sort_uniform(ctx.h, d, 0, n, 42))
    before = ctx.checksum(d, n, 8)
    ctx.ck(ctx.L.tg_radix_sort_local(ctx.h, C.byref(capi.u64_desc()), d, tmp, n))
    assert ctx.is_sorted(capi.u64_desc(), d, n)
    assert ctx.checksum(d, n, 8) == before
    # (sortedness + multiset equality is the full-size proof; the bit-exact comparison with the oracle / the reference's golden
    # outputs is made at 1e6 and 1e7 keys in test_gpu_sort_kernels.py)
    head = ctx.download(d, 4096 * 8, np.uint64)
    assert np.all(head[1:] >= head[:-1])
    ctx.free(d); ctx.free(tmp)


# ---- prefix sort: top digits by LSD passes + one finishing pass (tg_radix_sort.cu prefix_fixup_kernel) -----------

def test_prefix_sort_wide_keys_with_duplicates_stable(ctx):
    """16-byte items, wide u64 keys drawn from a pool (groups of ~6 fully equal keys): the finishing pass ranks
    equal keys by position, so the result is the stable sort — bit-exact with the oracle."""
    from thrill_b200 import capi
    n = 300000
    rng = np.random.RandomState(11)
    pool = rng.randint(0, 2**63 - 1, size=50000, dtype=np.int64).astype(np.uint64)
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = pool[rng.randint(0, len(pool), size=n)]
    kv["val"] = np.arange(n)
    before = ctx.L.tg_prefix_sort_fallbacks(ctx.h)
    out = _sort_on_gpu(ctx, kv, capi.kv_key_desc()).view(O.KV)
    assert ctx.L.tg_prefix_sort_fallbacks(ctx.h) == before          # the fast path handled it
    ref = O.sort_items(kv, O.KV_DESC).view(O.KV)
    assert np.array_equal(out, ref)


def test_prefix_sort_falls_back_on_long_equal_prefix_runs():
    """every distinct key 150 times (> the 64 items the finishing pass can see): the sort must notice and fall back
    to the plain LSD passes; the next sorts skip the attempt (penalty) and stay correct.  (Own ctx: the penalty
    counter of the shared one depends on which tests ran before.)"""
    from thrill_b200 import capi
    ctx = capi.Ctx(device=0)
    n = 300000
    rng = np.random.RandomState(12)
    pool = rng.randint(0, 2**63 - 1, size=2000, dtype=np.int64).astype(np.uint64)
    keys = np.repeat(pool, 150)
    rng.shuffle(keys)
    before = ctx.L.tg_prefix_sort_fallbacks(ctx.h)
    out = _sort_on_gpu(ctx, keys, capi.u64_desc()).view(np.uint64)
    assert np.array_equal(out, np.sort(keys))
    assert ctx.L.tg_prefix_sort_fallbacks(ctx.h) == before + 1
    for _ in range(10):         # penalty window, then a fresh (failing) attempt: always the right answer
        out = _sort_on_gpu(ctx, keys, capi.u64_desc()).view(np.uint64)
        assert np.array_equal(out, np.sort(keys))
    uni = O.gen_sort_uniform(3, 100000)
    for _ in range(10):
        assert np.array_equal(_sort_on_gpu(ctx, uni, capi.u64_desc()).view(np.uint64), np.sort(uni))
    ctx.close()


@pytest.mark.parametrize("n", [70000, 1 << 20])
def test_prefix_sort_groups_straddle_tiles(ctx, n):
    """keys = 40-bit random prefix in the high bits with only 2^14 distinct prefixes -> groups of 4..64 items that
    cross the finishing pass's 2048-item tiles; low bits random."""
    from thrill_b200 import capi
    rng = np.random.RandomState(n)
    npre = max(n // 24, 1)
    pre = rng.randint(0, 2**39, size=npre, dtype=np.int64).astype(np.uint64) << np.uint64(24)
    keys = pre[rng.randint(0, npre, size=n)] | rng.randint(0, 1 << 24, size=n).astype(np.uint64)
    out = _sort_on_gpu(ctx, keys, capi.u64_desc()).view(np.uint64)
    assert np.array_equal(out, np.sort(keys))


def test_prefix_sort_descending_and_big_endian(ctx):
    from thrill_b200 import capi
    keys = O.gen_sort_uniform(9, 400000)
    out = _sort_on_gpu(ctx, keys, capi.u64_desc(descending=True)).view(np.uint64)
    assert np.array_equal(out, np.sort(keys)[::-1])
    n = 250000
    rng = np.random.RandomState(5)
    t = np.zeros((n, 16), dtype=np.uint8)
    t[:, :10] = rng.randint(0, 256, size=(n, 10))
    t[:, 12:16] = np.arange(n, dtype=np.uint32).view(np.uint8).reshape(n, 4)
    t[1::2, :10] = t[0::2, :10]                         # every key twice: ties resolved by position (stable)
    desc = capi.KeyDesc(16, 0, 10, capi.KEY_BYTES_BE, 0, 0)
    out = _sort_on_gpu(ctx, t, desc).reshape(n, 16)
    ref = O.sort_items(t, O.KeyDesc(16, 0, 10, O.KEY_BYTES_BE)).reshape(n, 16)
    assert np.array_equal(out, ref)
