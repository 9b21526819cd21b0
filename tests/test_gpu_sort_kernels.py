"""GPU parity of the sample-sort kernels against the oracle (bit-exact): splitter classify+scatter
(api/sort.hpp:434-535), k-way merge (core/multiway_merge.hpp + tlx LoserTree), sampling, and the
single-GPU operator incl. the File <-> device codec.  pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden, sha
from gpu_util import make_blocks, u64p

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from thrill_b200 import capi
    c = capi.Ctx(device=0)
    yield c
    c.close()


def _oracle_classify_scatter(keys, gbase, samples_packed, p, desc=O.U64_DESC):
    spl = O.select_splitters(samples_packed, p, desc)
    padded, k = O.pad_splitters(spl, p, desc)
    tree = O.build_tree(padded, k, desc)
    b = O.classify(keys, gbase, tree, k, padded, desc).astype(np.int64)
    b[b == k - 1] = p - 1                                   # the writer swap, api/sort.hpp:460
    order = np.argsort(b, kind="stable")
    counts = np.bincount(b, minlength=p)
    return spl, order, counts


@pytest.mark.parametrize("p", [1, 2, 3, 5, 8, 16])
@pytest.mark.parametrize("dist", ["uniform", "dups"])
def test_classify_scatter_u64(ctx, p, dist):
    from thrill_b200 import capi
    n, gbase = 300001, 12345678
    rng = np.random.RandomState(p * 7 + len(dist))
    keys = O.gen_sort_uniform(5, n) if dist == "uniform" else rng.randint(0, 6, size=n).astype(np.uint64)
    sidx = np.sort(rng.choice(n, size=200, replace=False)).astype(np.uint64)
    samples = O.pack_samples(keys[sidx.astype(np.int64)], sidx + np.uint64(gbase))
    spl, order, counts = _oracle_classify_scatter(keys, gbase, samples, p)
    d_in = ctx.to_device(keys); d_out = ctx.alloc(n * 8)
    oc = np.zeros(p, dtype=np.uint64)
    spl_host = np.ascontiguousarray(spl) if p > 1 else np.zeros((1, 16), dtype=np.uint8)
    ctx.ck(ctx.L.tg_classify_scatter(ctx.h, C.byref(capi.u64_desc()), d_in, n, gbase, spl_host.ctypes.data, p, d_out, u64p(oc)))
    out = ctx.download(d_out, n * 8, np.uint64)
    assert np.array_equal(oc.astype(np.int64), counts)
    assert np.array_equal(out, keys[order])
    ctx.free(d_in); ctx.free(d_out)


def test_classify_scatter_pairs_stable(ctx):
    from thrill_b200 import capi
    n, p, gbase = 100000, 4, 0
    rng = np.random.RandomState(2)
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = rng.randint(0, 50, size=n); kv["val"] = np.arange(n)
    sidx = np.sort(rng.choice(n, size=97, replace=False)).astype(np.uint64)
    samples = O.pack_samples(kv[sidx.astype(np.int64)], sidx, O.KV_DESC)
    spl, order, counts = _oracle_classify_scatter(kv, gbase, samples, p, O.KV_DESC)
    d_in = ctx.to_device(kv); d_out = ctx.alloc(n * 16)
    oc = np.zeros(p, dtype=np.uint64)
    ctx.ck(ctx.L.tg_classify_scatter(ctx.h, C.byref(capi.kv_key_desc()), d_in, n, gbase, np.ascontiguousarray(spl).ctypes.data, p, d_out, u64p(oc)))
    out = ctx.download(d_out, n * 16, O.KV)
    assert np.array_equal(oc.astype(np.int64), counts)
    assert np.array_equal(out, kv[order])
    ctx.free(d_in); ctx.free(d_out)


def test_kway_merge_reference_kat(ctx):
    """tests/core/multiway_merge_test.cpp:33-86: 4 runs x 3 items from std::mt19937(0) % 100"""
    from thrill_b200 import capi
    vals = np.random.RandomState(0).randint(0, 2**32, size=12, dtype=np.uint64)
    elems = (vals % 100).astype(np.uint64)
    runs = [np.sort(elems[i * 3:(i + 1) * 3]) for i in range(4)]
    cat = np.concatenate(runs)
    d_runs = ctx.to_device(cat); d_out = ctx.alloc(256); d_tmp = ctx.alloc(256)
    ri = np.array([3, 3, 3, 3], dtype=np.uint64)
    ctx.ck(ctx.L.tg_kway_merge(ctx.h, C.byref(capi.u64_desc()), d_runs, u64p(ri), 4, d_out, d_tmp))
    out = ctx.download(d_out, 96, np.uint64)
    assert np.array_equal(out, np.sort(elems))
    assert np.array_equal(out, O.multiway_merge(runs).view(np.uint64))
    for d in (d_runs, d_out, d_tmp):
        ctx.free(d)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_kway_merge_u64_random_runs(ctx, k):
    from thrill_b200 import capi
    rng = np.random.RandomState(k)
    runs = [np.sort(rng.randint(0, 2**40, size=rng.randint(0, 300000)).astype(np.uint64)) for _ in range(k)]
    if k > 2:
        runs[1] = np.empty(0, dtype=np.uint64)
    cat = np.concatenate(runs) if sum(len(r) for r in runs) else np.empty(0, dtype=np.uint64)
    n = len(cat)
    d_runs = ctx.to_device(cat) if n else ctx.alloc(16)
    d_out = ctx.alloc(max(n * 8, 16)); d_tmp = ctx.alloc(max(n * 8, 16))
    ri = np.array([len(r) for r in runs], dtype=np.uint64)
    ctx.ck(ctx.L.tg_kway_merge(ctx.h, C.byref(capi.u64_desc()), d_runs, u64p(ri), k, d_out, d_tmp))
    out = ctx.download(d_out, n * 8, np.uint64)
    assert np.array_equal(out, O.multiway_merge(runs).view(np.uint64))
    for d in (d_runs, d_out, d_tmp):
        ctx.free(d)


def test_kway_merge_pairs_stable_ties_by_source(ctx):
    """equal keys come out in run order (stable LoserTree, loser_tree.hpp:246-272)"""
    from thrill_b200 import capi
    k = 5
    runs = []
    for s in range(k):
        kv = np.zeros(20000, dtype=O.KV)
        kv["key"] = np.sort(np.random.RandomState(s).randint(0, 100, size=20000))
        kv["val"] = s * 1000000 + np.arange(20000)
        runs.append(kv)
    cat = np.concatenate(runs)
    d_runs = ctx.to_device(cat); d_out = ctx.alloc(cat.nbytes); d_tmp = ctx.alloc(cat.nbytes)
    ri = np.array([len(r) for r in runs], dtype=np.uint64)
    ctx.ck(ctx.L.tg_kway_merge(ctx.h, C.byref(capi.kv_key_desc()), d_runs, u64p(ri), k, d_out, d_tmp))
    out = ctx.download(d_out, cat.nbytes, O.KV)
    ref = O.multiway_merge(runs, desc=O.KV_DESC, stable=True).view(O.KV)
    assert np.array_equal(out, ref)
    for d in (d_runs, d_out, d_tmp):
        ctx.free(d)


def test_draw_samples_are_items_with_their_global_index(ctx):
    from thrill_b200 import capi
    n, gbase = 1000000, 5 * 10**9
    keys = O.gen_sort_uniform(0, n)
    d_in = ctx.to_device(keys)
    ns = C.c_uint64()
    buf = np.zeros((O.sample_size(n), 16), dtype=np.uint8)
    ctx.ck(ctx.L.tg_draw_samples(ctx.h, C.byref(capi.u64_desc()), d_in, n, gbase, 99, buf.ctypes.data, C.byref(ns)))
    assert ns.value == min(n, O.sample_size(n)) == 1993
    k = buf[:, :8].copy().view(np.uint64).ravel(); gi = buf[:, 8:].copy().view(np.uint64).ravel()
    assert np.all(gi >= gbase) and np.all(gi < gbase + n)
    assert np.array_equal(k, keys[(gi - np.uint64(gbase)).astype(np.int64)])
    ctx.free(d_in)


def test_sort_operator_single_gpu_golden_and_file_codec(ctx):
    """nranks = 1 (cfg2 shape): tg_sort_file over 2 MiB-ish Blocks -> tg_fetch_output into BlockWriter geometry,
    compared with the output of the unmodified reference (tests/golden) and the oracle."""
    from thrill_b200 import capi
    g = golden()
    n = 1000000
    keys = O.gen_sort_uniform(0, n)
    blocks, nb, raw = make_blocks(capi, keys, 1 << 18)
    out_items = C.c_size_t()
    ctx.ck(ctx.L.tg_sort_file(ctx.h, C.byref(capi.u64_desc()), blocks, nb, 7, C.byref(out_items)))
    assert out_items.value == n
    cap = 64
    geo = (capi.BlockGeom * cap)()
    ng = ctx.L.tg_file_geometry(n, 8, 4096, 2 << 20, geo, cap)
    out = np.zeros(n, dtype=np.uint64)
    ob = (capi.Block * ng)()
    off = 0
    for i in range(ng):
        ob[i].data = out.ctypes.data + off
        ob[i].bytes = geo[i].bytes
        off += geo[i].bytes
    ctx.ck(ctx.L.tg_fetch_output(ctx.h, ob, ng))
    assert sha(out) == str(g["sort_uniform_1000000_w4_sha256"])
    assert np.array_equal(out, O.sort_items(keys).view(np.uint64))


def test_terasort_records_single_gpu_golden(ctx):
    """cfg4 shape (reduced): 100-byte records, 10-byte big-endian key, vs the unmodified reference's output"""
    from thrill_b200 import api
    g = golden()
    rec = O.gen_records(0, 20000)
    actx = api.Context.__new__(api.Context)
    actx._rank, actx._n, actx.tg, actx.rng_seed, actx._op_counter = 0, 1, ctx, 1, 0
    out = api.DIA(actx, rec).Sort().items
    assert out.shape == (20000, 100)
    assert np.array_equal(out[:8], g["terasort_20000_w3_head"])
    assert sha(out) == str(g["terasort_20000_w3_sha256"])
    # duplicates-heavy keys: stable order == the oracle's stable sort
    rec2 = O.gen_records(0, 50000)
    rec2[:, :9] = 0
    rec2[:, 9] = np.random.RandomState(3).randint(0, 4, size=50000)
    out2 = api.DIA(actx, rec2).Sort().items
    ref2 = O.sort_items(rec2, O.RECORD_DESC).reshape(-1, 100)
    assert np.array_equal(out2, ref2)


def test_r2_golden_terasort_1e6_and_sort_1e7(ctx):
    """larger reference outputs (tests/golden/reference_outputs_r2.npz): TeraSort 1e6 records, Sort 1e7 uniform keys and
    5e6 Zipf(1, 2^20) keys, single GPU, byte-identical to the unmodified reference"""
    from thrill_b200 import api
    from golden_util import golden_r2
    g = golden_r2()
    actx = api.Context.__new__(api.Context)
    actx._rank, actx._n, actx.tg, actx.rng_seed, actx._op_counter = 0, 1, ctx, 1, 0
    out = api.DIA(actx, O.gen_records(0, 1000000)).Sort().items
    assert np.array_equal(out[:4], g["terasort_1000000_head"]) and np.array_equal(out[-4:], g["terasort_1000000_tail"])
    assert sha(out) == str(g["terasort_1000000_sha256"])
    out = api.DIA(actx, O.gen_sort_uniform(0, 10000000)).Sort().items
    assert sha(out) == str(g["sort_uniform_10000000_sha256"])
    out = api.DIA(actx, O.gen_sort_zipf(0, 5000000, O.zipf_cdf(1 << 20))).Sort().items
    assert sha(out) == str(g["sort_zipf_u2^20_5000000_sha256"])


def test_terasort_full_size_properties(ctx):
    """cfg4 per-GPU share, reduced to 2e7 records (2 GB): device-generated records, sorted by the 10-byte key, multiset kept"""
    from thrill_b200 import capi
    n = 20000000
    d = ctx.alloc(n * 100)
    ctx.ck(ctx.L.tg_gen_records(ctx.h, d, 0, n, 42))
    before = ctx.checksum(d, n, 100)
    desc = capi.record_desc()
    op, on = C.c_void_p(), C.c_size_t()
    ctx.ck(ctx.L.tg_sort(ctx.h, C.byref(desc), d, n, 5, C.byref(op), C.byref(on)))
    assert on.value == n
    assert ctx.is_sorted(desc, op.value, n)
    assert ctx.checksum(op.value, n, 100) == before
    ctx.free(d)
