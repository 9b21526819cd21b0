"""GPU parity of the ReduceByKey kernels against the oracle and the golden outputs of the unmodified
reference: hash partition (bit-exact destinations, stable), hash aggregate (keys exact; values exact for
integer ops and the exact-mode doubles, rel 1e-9 for f64 sums — SURVEY.md §8d).  pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden
from gpu_util import make_blocks, u64p

pytestmark = pytest.mark.gpu
F64_RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from thrill_b200 import capi
    c = capi.Ctx(device=0)
    yield c
    c.close()


def _aggregate(ctx, kv, op):
    from thrill_b200 import capi
    n = len(kv)
    d_in = ctx.to_device(kv) if n else ctx.alloc(16)
    d_out = ctx.alloc(max(n * 16, 16) + 32)
    nd = C.c_uint64()
    ctx.ck(ctx.L.tg_hash_aggregate(ctx.h, C.byref(capi.KVDesc(16, op)), d_in, n, d_out, C.byref(nd)))
    out = ctx.download(d_out, nd.value * 16, O.KV)
    ctx.free(d_in); ctx.free(d_out)
    return np.sort(out, order="key")


@pytest.mark.parametrize("p", [1, 2, 3, 8, 13])
def test_hash_partition_bit_exact_and_stable(ctx, p):
    from thrill_b200 import capi
    n = 400003
    kv = O.gen_reduce_uniform(0, n, universe=1 << 20, exact=2)
    kv["key"][::1000] = 0                                              # the sentinel key hashes like any other
    dest = O.hash_partition_ids(kv["key"], p).astype(np.int64)         # Hash128to64(0,key) % p
    order = np.argsort(dest, kind="stable")
    d_in = ctx.to_device(kv); d_out = ctx.alloc(n * 16)
    oc = np.zeros(p, dtype=np.uint64)
    ctx.ck(ctx.L.tg_hash_partition(ctx.h, C.byref(capi.KVDesc(16, capi.OP_SUM_U64)), d_in, n, p, d_out, u64p(oc)))
    out = ctx.download(d_out, n * 16, O.KV)
    assert np.array_equal(oc.astype(np.int64), np.bincount(dest, minlength=p))
    assert np.array_equal(out, kv[order])
    ctx.free(d_in); ctx.free(d_out)


def test_hash_aggregate_reference_kats(ctx):
    """tests/core/reduce_hash_table_test.cpp:54-144 (500 keys incl. key 0) and
    tests/api/reduce_node_test.cpp:93-139 (1e6 pairs, 1000 keys)"""
    from thrill_b200 import capi
    i = np.arange(50000, dtype=np.uint64)
    kv = np.zeros(50000, dtype=O.KV); kv["key"] = i % 500; kv["val"] = i // 500
    out = _aggregate(ctx, kv, capi.OP_SUM_U64)
    assert np.array_equal(out["key"], np.arange(500, dtype=np.uint64)) and np.all(out["val"] == 100 * 99 // 2)
    i = np.arange(1000000, dtype=np.uint64)
    kv = np.zeros(1000000, dtype=O.KV); kv["key"] = i % 1000; kv["val"] = i // 1000
    out = _aggregate(ctx, kv, capi.OP_SUM_U64)
    assert len(out) == 1000 and np.all(out["val"] == 1000 * 999 // 2)


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 200000])
@pytest.mark.parametrize("op", ["sum_u64", "min_u64", "max_u64", "sum_f64_exact", "min_f64", "max_f64"])
def test_hash_aggregate_ops_vs_oracle(ctx, n, op):
    from thrill_b200 import capi
    rng = np.random.RandomState(n + len(op))
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = rng.randint(0, max(2, n // 7), size=n)                 # includes the zero key
    if op.endswith("u64"):
        kv["val"] = rng.randint(0, 1 << 40, size=n)
    elif op == "sum_f64_exact":
        kv["val"] = rng.randint(0, 1024, size=n).astype(np.float64).view(np.uint64)
    else:
        kv["val"] = (rng.rand(n) * 100 - 50).view(np.uint64)
    code = {"sum_u64": capi.OP_SUM_U64, "min_u64": capi.OP_MIN_U64, "max_u64": capi.OP_MAX_U64,
            "sum_f64_exact": capi.OP_SUM_F64, "min_f64": capi.OP_MIN_F64, "max_f64": capi.OP_MAX_F64}[op]
    ocode = {"sum_u64": O.OP_SUM_U64, "min_u64": O.OP_MIN_U64, "max_u64": O.OP_MAX_U64,
             "sum_f64_exact": O.OP_SUM_F64, "min_f64": O.OP_MIN_F64, "max_f64": O.OP_MAX_F64}[op]
    out = _aggregate(ctx, kv, code)
    ref = O.reduce_simple(kv, ocode) if n else np.empty(0, dtype=O.KV)
    assert np.array_equal(out, ref)


def test_reduce_operator_golden_zipf_f64(ctx):
    """cfg3 shape (reduced): ReducePair<u64,double>(plus) on Zipf keys vs the unmodified reference's output"""
    from thrill_b200 import capi
    g = golden()
    ref = g["reduce_f64_zipf_u4096_200000_w3"]
    kv = O.gen_reduce_zipf(0, 200000, O.zipf_cdf(4096))
    blocks, nb, raw = make_blocks(capi, kv, 1 << 20)
    n_out = C.c_size_t()
    ctx.ck(ctx.L.tg_reduce_file(ctx.h, C.byref(capi.KVDesc(16, capi.OP_SUM_F64)), blocks, nb, C.byref(n_out)))
    out = np.zeros(n_out.value, dtype=O.KV)
    ob = (capi.Block * 1)(); ob[0].data = out.ctypes.data; ob[0].bytes = out.nbytes
    ctx.ck(ctx.L.tg_fetch_output(ctx.h, ob, 1))
    out = np.sort(out, order="key")
    assert np.array_equal(out["key"], ref["key"])
    a, b = out["val"].view(np.float64), ref["val"].view(np.float64)
    assert np.all(np.abs(a - b) <= F64_RTOL * np.maximum(1.0, np.abs(b)))
    # exact mode: bit-identical sums
    ref = g["reduce_f64_exact_zipf_u4096_200000_w4"]
    kv = O.gen_reduce_zipf(0, 200000, O.zipf_cdf(4096), exact=1)
    out = _aggregate(ctx, kv, capi.OP_SUM_F64)
    assert np.array_equal(out["key"], ref["key"]) and np.array_equal(out["val"], ref["val"])
    ref = g["reduce_u64_uniform_u3000_100000_w2"]
    kv = O.gen_reduce_uniform(0, 100000, universe=3000, exact=2)
    out = _aggregate(ctx, kv, capi.OP_SUM_U64)
    assert np.array_equal(out["key"], ref["key"]) and np.array_equal(out["val"], ref["val"])


def test_reduce_large_device_generated_exact(ctx):
    """5e7 Zipf records generated on the device, exact-mode values: every sum is bit-exact against the
    oracle's straightforward aggregate, and the number of records is conserved (sum of counts)."""
    from thrill_b200 import capi
    n, U = 50000000, 1 << 22
    cdf = O.zipf_cdf(U)
    d_cdf = ctx.to_device(cdf)
    d_in = ctx.alloc(n * 16)
    ctx.ck(ctx.L.tg_gen_reduce_zipf(ctx.h, d_in, 0, n, 42, d_cdf, U, 1))
    out_p = C.c_void_p(); out_n = C.c_size_t()
    ctx.ck(ctx.L.tg_reduce_by_key(ctx.h, C.byref(capi.KVDesc(16, capi.OP_SUM_F64)), d_in, n, C.byref(out_p), C.byref(out_n)))
    out = np.sort(ctx.download(out_p.value, out_n.value * 16, O.KV), order="key")
    kv = O.gen_reduce_zipf(0, n, cdf, exact=1)
    ref = O.reduce_simple(kv, O.OP_SUM_F64)
    assert np.array_equal(out, ref)
    ctx.free(d_in); ctx.free(d_cdf)


@pytest.mark.parametrize("op", ["sum_u64", "min_u64", "max_u64", "sum_f64_exact", "min_f64", "max_f64"])
def test_partitioned_aggregate_ops_hot_and_zero_keys(ctx, op):
    """n above the threshold of the partitioned aggregation (two hash-digit passes + shared-memory tables): a hot key
    whose segment is cut into several units (merged afterwards), the zero key (side slot), and a long tail."""
    from thrill_b200 import capi
    n = 700001
    rng = np.random.RandomState(len(op))
    kv = np.zeros(n, dtype=O.KV)
    keys = rng.randint(1, 100000, size=n).astype(np.uint64)
    sel = rng.rand(n)
    keys[sel < 0.30] = 7                    # hot: ~210000 records = ~100 units
    keys[(sel >= 0.30) & (sel < 0.40)] = 0  # the sentinel key
    keys[(sel >= 0.40) & (sel < 0.45)] = 2**63 + 12345
    kv["key"] = keys
    if op.endswith("u64"):
        kv["val"] = rng.randint(0, 1 << 40, size=n)
    elif op == "sum_f64_exact":
        kv["val"] = rng.randint(0, 1024, size=n).astype(np.float64).view(np.uint64)
    else:
        kv["val"] = (rng.rand(n) * 100 - 50).view(np.uint64)
    code = {"sum_u64": capi.OP_SUM_U64, "min_u64": capi.OP_MIN_U64, "max_u64": capi.OP_MAX_U64,
            "sum_f64_exact": capi.OP_SUM_F64, "min_f64": capi.OP_MIN_F64, "max_f64": capi.OP_MAX_F64}[op]
    ocode = {"sum_u64": O.OP_SUM_U64, "min_u64": O.OP_MIN_U64, "max_u64": O.OP_MAX_U64,
             "sum_f64_exact": O.OP_SUM_F64, "min_f64": O.OP_MIN_F64, "max_f64": O.OP_MAX_F64}[op]
    out = _aggregate(ctx, kv, code)
    ref = O.reduce_simple(kv, ocode)
    assert np.array_equal(out, ref)


def test_partitioned_aggregate_all_distinct_and_f64_tolerance(ctx):
    """every key once (nothing to reduce, every segment full of distinct keys), then uniform duplicates with real
    doubles: keys exact, sums within the stated tolerance"""
    from thrill_b200 import capi
    n = 1 << 20
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = np.random.RandomState(3).permutation(n).astype(np.uint64) * np.uint64(2654435761) + np.uint64(1)
    kv["val"] = np.arange(n, dtype=np.uint64)
    out = _aggregate(ctx, kv, capi.OP_SUM_U64)
    assert np.array_equal(out, np.sort(kv, order="key"))
    kv = O.gen_reduce_uniform(0, 900000, universe=1 << 17)
    out = _aggregate(ctx, kv, capi.OP_SUM_F64)
    ref = O.reduce_simple(kv, O.OP_SUM_F64)
    assert np.array_equal(out["key"], ref["key"])
    a, b = out["val"].view(np.float64), ref["val"].view(np.float64)
    assert np.all(np.abs(a - b) <= F64_RTOL * np.maximum(1.0, np.abs(b)))


# ---- ReduceToIndex (api/reduce_to_index.hpp:60-237, the PageRank step) ----------------------------------------------------

def _reduce_to_index(ctx, kv, size, op, neutral=(0, 0)):
    from thrill_b200 import capi
    n = len(kv)
    d_in = ctx.to_device(kv) if n else ctx.alloc(16)
    neu = np.zeros(1, dtype=O.KV); neu["key"], neu["val"] = neutral
    op_, on, ob = C.c_void_p(), C.c_size_t(), C.c_uint64()
    ctx.ck(ctx.L.tg_reduce_to_index(ctx.h, C.byref(capi.KVDesc(16, op)), d_in, n, size, neu.ctypes.data,
                                    C.byref(op_), C.byref(on), C.byref(ob)))
    out = ctx.download(op_.value, on.value * 16, O.KV)
    ctx.free(d_in)
    return out, ob.value


def test_reduce_to_index_golden_and_oracle(ctx):
    """dense result vs the unmodified reference (golden) — exact-mode doubles bit-identical, missing indices neutral — and
    vs the oracle for the other reduce functions, a non-zero neutral element and an input above the partitioned-aggregation
    threshold"""
    from thrill_b200 import capi
    g = golden()
    kv = O.gen_reduce_zipf(0, 20000, O.zipf_cdf(1000), exact=1); kv["key"] %= 1000
    out, begin = _reduce_to_index(ctx, kv, 1000, capi.OP_SUM_F64)
    assert begin == 0 and np.array_equal(out, g["reduce_to_index_zipf_u1000_20000_exact1_w3"])
    kv = O.gen_reduce_uniform(0, 3000, universe=50000, exact=1); kv["key"] %= 50000
    out, _ = _reduce_to_index(ctx, kv, 50000, capi.OP_SUM_F64)
    assert np.array_equal(out, g["reduce_to_index_uniform_u50000_3000_exact1_w5"])
    kv = O.gen_reduce_zipf(0, 20000, O.zipf_cdf(1000)); kv["key"] %= 1000
    out, _ = _reduce_to_index(ctx, kv, 1000, capi.OP_SUM_F64)
    ref = g["reduce_to_index_zipf_u1000_20000_exact0_w4"]
    assert np.array_equal(out["key"], ref["key"])
    a, b = out["val"].view(np.float64), ref["val"].view(np.float64)
    assert np.all(np.abs(a - b) <= F64_RTOL * np.maximum(1.0, np.abs(b)))
    rng = np.random.RandomState(8)
    n, size = 900000, 300000
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = rng.randint(0, size, size=n) // 3 * 3            # two thirds of the indices stay neutral
    kv["val"] = rng.randint(0, 1 << 30, size=n)
    for op, oop in ((capi.OP_SUM_U64, O.OP_SUM_U64), (capi.OP_MAX_U64, O.OP_MAX_U64), (capi.OP_MIN_U64, O.OP_MIN_U64)):
        out, _ = _reduce_to_index(ctx, kv, size, op, neutral=(7, 99))
        assert np.array_equal(out, O.reduce_to_index(kv, size, oop, neutral=(7, 99)))


def test_reduce_to_index_rejects_out_of_range_index(ctx):
    from thrill_b200 import capi
    kv = np.zeros(100, dtype=O.KV)
    kv["key"] = np.arange(100); kv["key"][17] = 5000
    with pytest.raises(capi.ThrillGpuError):
        _reduce_to_index(ctx, kv, 1000, capi.OP_SUM_U64)


def test_r2_golden_reduce_4e6_zipf_exact(ctx):
    """ReducePair of 4e6 Zipf(1, 2^20) records in exact mode: digest of the key-sorted result equals the unmodified
    reference's (tests/golden/reference_outputs_r2.npz); exercises the hot-key folding of the counting read"""
    from thrill_b200 import capi
    from golden_util import golden_r2, sha
    g = golden_r2()
    kv = O.gen_reduce_zipf(0, 4000000, O.zipf_cdf(1 << 20), exact=1)
    out = _aggregate(ctx, kv, capi.OP_SUM_F64)
    assert len(out) == int(g["reduce_f64_exact_zipf_u2^20_4000000_distinct"])
    assert sha(out) == str(g["reduce_f64_exact_zipf_u2^20_4000000_sha256"])


@pytest.mark.parametrize("universe", [100, 700])
def test_many_very_frequent_keys(ctx, universe):
    """a small universe: every key is seen hundreds of times in the 65536-record sample, more keys than there are warp-private
    accumulators (64) — all of them are folded by the counting read, some in the CTA-shared table; exact sums"""
    from thrill_b200 import capi
    n = 3000000
    kv = O.gen_reduce_uniform(0, n, universe=universe, exact=1)
    kv["key"][::1000] = 0                                    # the zero key (side slot) among them
    out = _aggregate(ctx, kv, capi.OP_SUM_F64)
    assert np.array_equal(out, O.reduce_simple(kv, O.OP_SUM_F64))
    out = _aggregate(ctx, kv, capi.OP_MAX_F64)
    assert np.array_equal(out, O.reduce_simple(kv, O.OP_MAX_F64))
