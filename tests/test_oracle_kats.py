"""Pins oracle/thrill_oracle.c against the known-answer tests of the reference's own suite
(SURVEY.md §8c), restated with the same inputs.  CPU only."""
import numpy as np
import pytest

import oracle_lib as O


# ---- generators / hashing ---------------------------------------------------
def test_splitmix64_known_values():
    # splitmix64 reference outputs (Vigna's splitmix64.c seeded with 0: first three next() values)
    # next() = mix(state += GOLDEN); our to_splitmix64(x) = mix(x + GOLDEN)
    G = 0x9E3779B97F4A7C15
    assert O.splitmix64(0) == 0xE220A8397B1DCDAF
    assert O.splitmix64(G) == 0x6E789E6AA1B965F4
    assert O.splitmix64((2 * G) % 2**64) == 0x06C45D188009454F


def _hash128to64_py(upper, lower):
    # common/hash.hpp:64-73
    M = 2**64 - 1
    k = 0x9DDFEA08EB382D69
    a = ((lower ^ upper) * k) & M
    a ^= a >> 47
    b = ((upper ^ a) * k) & M
    b ^= b >> 47
    return (b * k) & M


def test_hash128to64_matches_python_restatement():
    rng = np.random.RandomState(1)
    for _ in range(200):
        u, l = int(rng.randint(0, 2**63)) * 2 + 1, int(rng.randint(0, 2**63))
        assert O.hash128to64(u, l) == _hash128to64_py(u, l)
    assert O.hash128to64(0, 0) == 0
    keys = np.arange(1, 1000, dtype=np.uint64)
    pid = O.hash_partition_ids(keys, 13)
    assert all(int(pid[i]) == _hash128to64_py(0, int(keys[i])) % 13 for i in range(len(keys)))


def test_sample_size_formula():
    # common/reservoir_sampling.hpp:270-275, eps = 0.1: floor(log2(n) * 100), min 1
    assert O.sample_size(1) == 1
    assert O.sample_size(2) == 99             # 1/(0.1*0.1) = 99.99999999999999 in double, as in the reference
    assert O.sample_size(10**8) == 2657
    assert O.sample_size(125000000) == 2689


# ---- k-way merge: tests/core/multiway_merge_test.cpp:33-86 ------------------
def _mt19937_0():
    rs = np.random.RandomState(0)          # init_genrand(0) == std::mt19937(0)
    vals = rs.randint(0, 2**32, size=12, dtype=np.uint64)
    return vals


def test_multiway_merge_reference_kat():
    vals = _mt19937_0()
    assert int(vals[0]) == 2357136044      # first std::mt19937(0) output
    a, b = 4, 3
    elems = (vals % 100).astype(np.uint64)
    runs = [np.sort(elems[i * b:(i + 1) * b]) for i in range(a)]
    out = O.multiway_merge(runs).view(np.uint64)
    assert np.array_equal(out, np.sort(elems))
    out_s = O.multiway_merge(runs, stable=True).view(np.uint64)
    assert np.array_equal(out_s, np.sort(elems))


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 13])
@pytest.mark.parametrize("stable", [False, True])
def test_multiway_merge_random_and_empty_runs(k, stable):
    rng = np.random.RandomState(k)
    runs = [np.sort(rng.randint(0, 50, size=rng.randint(0, 200)).astype(np.uint64)) for _ in range(k)]
    runs[k // 2] = np.empty(0, dtype=np.uint64)
    out = O.multiway_merge(runs, stable=stable).view(np.uint64)
    assert np.array_equal(out, np.sort(np.concatenate(runs)))


def test_multiway_merge_stable_ties_by_source():
    # stable LoserTree: equal keys come out in source order (loser_tree.hpp:246-272)
    k = 5
    runs = []
    for s in range(k):
        kv = np.zeros(40, dtype=O.KV)
        kv["key"] = np.repeat(np.arange(10, dtype=np.uint64), 4)
        kv["val"] = s * 1000 + np.arange(40)
        runs.append(kv)
    out = O.multiway_merge(runs, desc=O.KV_DESC, stable=True).view(O.KV)
    for key in range(10):
        vals = out["val"][out["key"] == key]
        assert np.array_equal(vals, np.sort(vals))    # source-major, then position


# ---- probing table: tests/core/reduce_hash_table_test.cpp:54-144 ------------
def test_probing_table_kat_500_keys():
    test_size, mod = 50000, 500
    kv = np.zeros(test_size, dtype=O.KV)
    i = np.arange(test_size, dtype=np.uint64)
    kv["key"] = i % mod            # includes key 0 -> sentinel slot path
    kv["val"] = i // mod
    out, part = O.reduce_pre_phase(kv, 13, 1024 * 1024, O.OP_SUM_U64)
    out = np.sort(out, order="key")
    assert len(out) == mod
    assert np.array_equal(out["key"], np.arange(mod, dtype=np.uint64))
    assert np.all(out["val"] == (test_size // mod) * (test_size // mod - 1) // 2)


# ---- pre phase: tests/core/reduce_pre_phase_test.cpp:44-127 -----------------
def test_pre_phase_kat_601_keys_and_partitioning():
    mod = 601
    test_size = mod * 100
    kv = np.zeros(test_size, dtype=O.KV)
    i = np.arange(test_size, dtype=np.uint64)
    kv["key"] = i % mod
    kv["val"] = i // mod
    out, part = O.reduce_pre_phase(kv, 13, 1024 * 1024, O.OP_SUM_U64)
    # every emitted item sits in partition Hash128to64(0,key) % 13 (core/reduce_functional.hpp:60-72)
    assert np.array_equal(part, O.hash_partition_ids(out["key"], 13))
    out = np.sort(out, order="key")
    assert len(out) == mod
    assert np.all(out["val"] == 100 * 99 // 2)


# ---- post phase w/ spill + re-reduce: tests/core/reduce_post_phase_test.cpp:36-112
def test_post_phase_kat_spills_with_64k():
    mod = 601
    test_size = mod * 100
    kv = np.zeros(test_size, dtype=O.KV)
    i = np.arange(test_size, dtype=np.uint64)
    kv["key"] = i % mod
    kv["val"] = i // mod
    out, iters = O.reduce_post_phase(kv, 64 * 1024, O.OP_SUM_U64)
    out = np.sort(out, order="key")
    assert len(out) == mod
    assert np.array_equal(out["key"], np.arange(mod, dtype=np.uint64))
    assert np.all(out["val"] == 100 * 99 // 2)


def test_post_phase_tiny_memory_forces_salted_rereduce():
    rng = np.random.RandomState(3)
    kv = np.zeros(200000, dtype=O.KV)
    kv["key"] = rng.randint(0, 20000, size=len(kv))
    kv["val"] = 1
    out, iters = O.reduce_post_phase(kv, 32 * 1024, O.OP_SUM_U64)
    assert iters >= 1                                   # spilled partitions were re-reduced
    ref = O.reduce_simple(kv, O.OP_SUM_U64)
    assert np.array_equal(np.sort(out, order="key"), ref)


# ---- ReduceNode end to end: tests/api/reduce_node_test.cpp:47-139 -----------
@pytest.mark.parametrize("p", [1, 2, 3, 5, 8])
def test_reduce_operator_mod4_sums(p):
    v = np.arange(1, 17, dtype=np.uint64)
    kv = np.zeros(16, dtype=O.KV)
    kv["key"] = v % 4
    kv["val"] = v
    counts = np.full(p, 16 // p, dtype=np.uint64); counts[-1] += 16 - counts.sum()
    out, oc = O.reduce_operator(kv, counts, 1 << 20, O.OP_SUM_U64)
    assert sorted(out["val"].tolist()) == [28, 32, 36, 40]


@pytest.mark.parametrize("p", [1, 2, 5, 8])
def test_reduce_operator_pairs_1000_keys(p):
    test_size, mod = 1000000, 1000
    i = np.arange(test_size, dtype=np.uint64)
    kv = np.zeros(test_size, dtype=O.KV)
    kv["key"] = i % mod
    kv["val"] = i // mod
    counts = np.full(p, test_size // p, dtype=np.uint64); counts[-1] += test_size - counts.sum()
    out, oc = O.reduce_operator(kv, counts, 8 << 20, O.OP_SUM_U64)
    assert len(out) == mod
    div = test_size // mod
    assert np.all(out["val"] == div * (div - 1) // 2)
    # each key lives on worker Hash128to64(0,key) % p
    owner = np.repeat(np.arange(p), oc.astype(np.int64))
    assert np.array_equal(owner, O.hash_partition_ids(out["key"], p))


# ---- SortNode: tests/api/sort_node_test.cpp ---------------------------------
def _even_counts(n, p):
    c = np.full(p, n // p, dtype=np.uint64)
    c[: n % p] += 1
    return c


@pytest.mark.parametrize("p", [1, 2, 3, 5, 8])
def test_sort_operator_known_integers(p):            # :25-53 (size reduced from 6e6)
    n = 600000
    keys = (n - np.arange(n) - 1).astype(np.uint64)
    out, oc = O.sort_operator(keys, _even_counts(n, p))
    assert np.array_equal(out.view(np.uint64), np.arange(n, dtype=np.uint64))
    assert int(oc.sum()) == n


@pytest.mark.parametrize("p", [1, 2, 5, 8])
def test_sort_operator_all_equal_and_four_values(p):   # :162-185, :187-210
    ones = np.ones(10000, dtype=np.uint64)
    out, oc = O.sort_operator(ones, _even_counts(10000, p))
    assert np.all(out.view(np.uint64) == 1)
    # the global-index tie-break (api/sort.hpp:487-502) balances all-equal input across workers
    if p > 1:
        assert oc.max() <= 10000 * 2 // p + 200
    z3 = (np.arange(10000) % 4).astype(np.uint64)
    out, oc = O.sort_operator(z3, _even_counts(10000, p))
    assert np.array_equal(out.view(np.uint64), (np.arange(10000) * 4 // 10000).astype(np.uint64))


@pytest.mark.parametrize("p", [1, 2, 5, 8])
def test_sort_operator_one_zero_and_empty_workers(p):  # :212-276
    out, oc = O.sort_operator(np.zeros(1, dtype=np.uint64), np.array([1] + [0] * (p - 1), dtype=np.uint64))
    assert out.view(np.uint64).tolist() == [0] and int(oc.sum()) == 1
    out, oc = O.sort_operator(np.empty(0, dtype=np.uint64), np.zeros(p, dtype=np.uint64))
    assert len(out) == 0 and int(oc.sum()) == 0
    rng = np.random.RandomState(7)
    keys = rng.randint(0, 2**62, size=10000).astype(np.uint64)
    counts = np.zeros(p, dtype=np.uint64); counts[0] = 10000        # all items on worker 0
    out, oc = O.sort_operator(keys, counts)
    assert np.array_equal(out.view(np.uint64), np.sort(keys))


@pytest.mark.parametrize("p", [2, 3, 8])
def test_sort_operator_stable_pairs(p):                # :291-404 (index increasing within equal values)
    n = 50000
    rng = np.random.RandomState(5)
    kv = np.zeros(n, dtype=O.KV)
    kv["key"] = rng.randint(0, 100, size=n)
    kv["val"] = np.arange(n)
    out, oc = O.sort_operator(kv, _even_counts(n, p), desc=O.KV_DESC, stable=True)
    out = out.view(O.KV)
    assert np.all(np.diff(out["key"].astype(np.int64)) >= 0)
    same = out["key"][1:] == out["key"][:-1]
    assert np.all(out["val"][1:][same] > out["val"][:-1][same])


def test_classify_tie_break_matches_python_restatement():
    """TransmitItems descent + EqualSampleGreaterIndex (api/sort.hpp:478-502) in straight Python."""
    rng = np.random.RandomState(11)
    p = 5
    n = 4000
    keys = rng.randint(0, 6, size=n).astype(np.uint64)          # massive duplicates
    sidx = np.sort(rng.choice(n, size=60, replace=False))
    samples = O.pack_samples(keys[sidx], sidx.astype(np.uint64))
    spl = O.select_splitters(samples, p)
    padded, k = O.pad_splitters(spl, p)
    tree = O.build_tree(padded, k)
    got = O.classify(keys, 0, tree, k, padded)
    sk = padded[:k - 1, :8].copy().view(np.uint64).ravel()
    si = padded[:k - 1, 8:].copy().view(np.uint64).ravel()
    tr = tree.view(np.uint64).ravel()
    log_k = k.bit_length() - 1
    for i in range(n):
        j = 1
        for _ in range(log_k):
            j = 2 * j + (0 if keys[i] < tr[j] else 1)
        b = j - k
        while b and (not sk[b - 1] < keys[i]) and si[b - 1] >= i:
            b -= 1
        assert got[i] == b
    # buckets are monotone in (key, index) order
    order = np.lexsort((np.arange(n), keys))
    assert np.all(np.diff(got[order].astype(np.int64)) >= 0)


# ---- File / Block layout: tests/data/file_test.cpp:30-122 --------------------
def test_file_layout_block_sizes_and_item_starts():
    # 94-byte frozen image uses 16-byte blocks: 6 blocks of 16,16,16,16,16,14 (file_test.cpp:50-60).
    # Same geometry with 47 fixed 2-byte items:
    m = O.file_layout(47, 2, start_block_size=4096, max_block_size=16)
    assert [int(x) for x in m["end"]] == [16, 16, 16, 16, 16, 14]
    assert int(m["num_items"].sum()) == 47
    # straddling 100-byte records over the doubling 4 KiB.. blocks (block_writer.hpp:405-420)
    m = O.file_layout(1000, 100)
    assert int(m["end"].sum()) == 100000
    assert int(m["num_items"].sum()) == 1000
    assert [int(x) for x in m["end"][:4]] == [4096, 8192, 16384, 32768]
    pos = 0
    for blk in m:
        first_start = (-(-pos // 100)) * 100             # first item start at/after block begin
        assert int(blk["first_item"]) == first_start - pos
        assert int(blk["num_items"]) == len(range(first_start, pos + int(blk["end"]), 100))
        pos += int(blk["end"])
    # default writer never exceeds 1 MiB blocks: 2*bs < 2 MiB stops the doubling at 1 MiB
    m = O.file_layout(1 << 20, 8)
    assert int(m["end"].max()) == 1 << 20
