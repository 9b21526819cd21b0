"""world_size-2 (and 3) CPU tests over gloo of the host logic on the N>1 path: the control-plane arithmetic the
GPU operator performs around its kernels (shard ranges, sample counts, identical splitter selection on every
rank from all-gathered samples, count exchange -> Alltoallv offsets) with the per-rank compute done by the
oracle (test infrastructure).  The distributed result must equal the output of the unmodified reference
(tests/golden).  Also covers bench.py's cross-rank reductions."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _all_gather(arr, world):
    parts = [None] * world
    dist.all_gather_object(parts, arr)
    return parts


def _sort_worker(rank, world, port, n, out_dir):
    _init(rank, world, port)
    import oracle_lib as O
    from thrill_b200 import api, capi
    L = capi.lib()
    # shard exactly as api.Generate / common::CalculateLocalRange does
    lo, hi = api._local_range(n, world, rank)
    keys = O.gen_sort_uniform(lo, hi - lo)
    n_local = hi - lo
    # (1) ExPrefixSumTotal
    counts = [int(x[0]) for x in _all_gather(np.array([n_local], dtype=np.uint64), world)]
    prefix = sum(counts[:rank])
    assert prefix == lo
    # (2) samples (item, global index), all-gathered; every rank selects the same splitters with the C ABI
    ns = min(n_local, L.tg_sample_size(n_local))
    rng = np.random.RandomState(1000 + rank)
    idx = rng.randint(0, n_local, size=ns)
    mine = O.pack_samples(keys[idx], (idx + prefix).astype(np.uint64))
    allsamp = np.ascontiguousarray(np.concatenate(_all_gather(mine, world)))
    spl = np.zeros((world - 1, 16), dtype=np.uint8)
    d = capi.u64_desc()
    assert L.tg_select_splitters(C.byref(d), allsamp.ctypes.data, len(allsamp), world, spl.ctypes.data) == 0
    all_spl = _all_gather(spl, world)
    assert all(np.array_equal(all_spl[0], s) for s in all_spl)          # identical on every rank
    # (3) classify (oracle = checker standing in for the CUDA kernel), count exchange, Alltoallv
    padded, k = O.pad_splitters(spl, world)
    tree = O.build_tree(padded, k)
    b = O.classify(keys, prefix, tree, k, padded).astype(np.int64)
    b[b == k - 1] = world - 1
    send = [np.ascontiguousarray(keys[b == r]) for r in range(world)]
    send_cnt = np.array([len(s) for s in send], dtype=np.int64)
    mat = np.stack(_all_gather(send_cnt, world))                        # [src][dst]
    recv_cnt = mat[:, rank]
    # gloo has no all_to_all: exchange through all_gather_object (test-scale data)
    allsend = _all_gather(send, world)
    got = [allsend[src][rank] for src in range(world)]
    assert [len(g) for g in got] == [int(c) for c in recv_cnt]
    cat = np.concatenate(got)
    local_sorted = O.sort_items(cat).view(np.uint64) if len(cat) else np.empty(0, np.uint64)
    np.save(os.path.join(out_dir, "part%d.npy" % rank), local_sorted)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_sample_sort_host_logic_matches_reference(tmp_path, world):
    from golden_util import golden
    n = 4096
    port = 29511 + world
    mp.spawn(_sort_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    out = np.concatenate([np.load(os.path.join(str(tmp_path), "part%d.npy" % r)) for r in range(world)])
    assert np.array_equal(out, golden()["sort_uniform_4096_w3"])          # same generator, same global result


def _reduce_worker(rank, world, port, out_dir):
    _init(rank, world, port)
    import oracle_lib as O
    from thrill_b200 import api
    n = 200000
    lo, hi = api._local_range(n, world, rank)
    cdf = O.zipf_cdf(4096)
    kv = O.gen_reduce_zipf(lo, hi - lo, cdf, exact=1)
    pre, part = O.reduce_pre_phase(kv, world, 32 << 20, O.OP_SUM_F64)       # partition = Hash128to64(0,key) % p
    send = [np.ascontiguousarray(pre[part == r]) for r in range(world)]
    allsend = _all_gather(send, world)
    got = np.concatenate([allsend[src][rank] for src in range(world)])
    out, _ = O.reduce_post_phase(got, 32 << 20, O.OP_SUM_F64)
    np.save(os.path.join(out_dir, "red%d.npy" % rank), out)
    # bench.py's cross-rank reductions over gloo
    import bench
    assert bench.max_over_ranks(float(rank + 1), world) == float(world)
    assert bench.gather_objects(rank, world) == list(range(world))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_reduce_host_logic_matches_reference(tmp_path):
    import oracle_lib as O
    from golden_util import golden
    world = 2
    mp.spawn(_reduce_worker, args=(world, 29530, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "red%d.npy" % r)) for r in range(world)]
    for r, p in enumerate(parts):                                          # ownership: key -> hash % p
        assert np.all(O.hash_partition_ids(p["key"], world) == r)
    out = np.sort(np.concatenate(parts), order="key")
    ref = golden()["reduce_f64_exact_zipf_u4096_200000_w4"]               # exact mode: independent of p
    assert np.array_equal(out["key"], ref["key"]) and np.array_equal(out["val"], ref["val"])


def _rti_worker(rank, world, port, out_dir):
    """ReduceToIndex around its kernels: range partition k*p/size (common/math.hpp:98-100), the worker's index range
    Range(0,size).Partition(r,p) (:85-94), exchange, dense post phase — per-rank compute by the oracle"""
    _init(rank, world, port)
    import oracle_lib as O
    from thrill_b200 import api
    n, size = 20000, 1000
    lo, hi = api._local_range(n, world, rank)
    kv = O.gen_reduce_zipf(lo, hi - lo, O.zipf_cdf(size), exact=1)
    kv["key"] %= size
    pre = O.reduce_simple(kv, O.OP_SUM_F64)                                  # pre phase: local aggregation
    dest = (pre["key"] * np.uint64(world)) // np.uint64(size)
    send = [np.ascontiguousarray(pre[dest == r]) for r in range(world)]
    allsend = _all_gather(send, world)
    got = np.concatenate([allsend[src][rank] for src in range(world)])
    begin, end = (rank * size + world - 1) // world, ((rank + 1) * size + world - 1) // world
    assert len(got) == 0 or (got["key"].min() >= begin and got["key"].max() < end)
    got = got.copy(); got["key"] -= np.uint64(begin)
    dense = O.reduce_to_index(got, end - begin, O.OP_SUM_F64)
    present = ~((dense["key"] == 0) & (dense["val"] == 0))
    dense["key"][present] += np.uint64(begin)
    if begin == 0 and len(got) and got["key"].min() == 0:
        pass                                                                  # index 0 keeps key 0
    np.save(os.path.join(out_dir, "rti%d.npy" % rank), dense)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_reduce_to_index_host_logic_matches_reference(tmp_path, world):
    sys.path.insert(0, HERE)
    from golden_util import golden
    mp.spawn(_rti_worker, args=(world, 29720 + world, str(tmp_path)), nprocs=world, join=True)
    cat = np.concatenate([np.load(os.path.join(str(tmp_path), "rti%d.npy" % r)) for r in range(world)])
    assert np.array_equal(cat, golden()["reduce_to_index_zipf_u1000_20000_exact1_w3"])


def test_local_range_matches_thrill_generate_split():
    from thrill_b200 import api
    for n in [0, 1, 7, 100, 4096, 10**8]:
        for p in [1, 2, 3, 5, 8]:
            r = [api._local_range(n, p, i) for i in range(p)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(p - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_bench_zipf_table_equals_oracle_table():
    import bench
    import oracle_lib as O
    assert np.array_equal(bench.zipf_cdf_numpy(4096), O.zipf_cdf(4096))
    assert np.array_equal(bench.zipf_cdf_numpy(1 << 16), O.zipf_cdf(1 << 16))


def test_bench_parity_helpers_match_the_oracle():
    """bench.py's parity_check uses its own numpy restatement of Hash128to64 (it may not import the oracle on that path):
    pinned here against the C oracle, which is pinned against the reference"""
    import bench
    import oracle_lib as O
    keys = np.concatenate([np.arange(0, 5000, dtype=np.uint64), O.gen_sort_uniform(0, 20000)])
    for p in (1, 2, 3, 8, 13):
        assert np.array_equal(bench.hash128to64_np(keys) % np.uint64(p), O.hash_partition_ids(keys, p).astype(np.uint64))
    assert np.array_equal(bench.zipf_cdf_numpy(4096), O.zipf_cdf(4096))


def _plan_worker(rank, world, port, out_dir):
    """every rank computes its part of the exchange plan with the PRODUCT's arithmetic (tg_exchange_plan of libthrill_gpu.so,
    pure host code) from the all-gathered count matrix, as the GPU operators do after their ncclAllGather"""
    import ctypes as C
    import torch.distributed as dist
    from thrill_b200 import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(100 + rank)
    mine = rs.randint(0, 1000, size=world).astype(np.uint32)           # items this rank holds for every destination
    if rank == 1:
        mine[0] = 0                                                     # an empty (src, dst) pair
    rows = [None] * world
    dist.all_gather_object(rows, mine)
    mat = np.ascontiguousarray(np.stack(rows), dtype=np.uint32)
    L = capi.lib()
    send = (C.c_uint64 * world)(); recv = (C.c_uint64 * world)(); before = (C.c_uint64 * world)()
    n_recv, worst = C.c_uint64(), C.c_uint64()
    assert L.tg_exchange_plan(world, rank, mat.ctypes.data_as(C.POINTER(C.c_uint32)), send, recv, before, C.byref(n_recv), C.byref(worst)) == 0
    np.save(os.path.join(out_dir, "plan%d.npy" % rank),
            np.array([list(send), list(recv), list(before), [n_recv.value] * world, [worst.value] * world], dtype=np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_plan_is_consistent_across_ranks(tmp_path, world):
    """N > 1 host logic of the fused exchange on CPU (gloo): what rank s sends to d is what d expects from s, every window is tiled
    without gaps or overlaps in rank order, and all ranks agree on the largest receive size (the uniform growth / error verdict)"""
    mp.spawn(_plan_worker, args=(world, 29540 + world, str(tmp_path)), nprocs=world, join=True)
    plans = [np.load(os.path.join(str(tmp_path), "plan%d.npy" % r)) for r in range(world)]
    for s in range(world):
        for d in range(world):
            assert plans[s][0][d] == plans[d][1][s]                     # send_cnt[s -> d] == recv_cnt[d <- s]
    for d in range(world):
        off = 0
        for s in range(world):                                          # shares lie back to back in rank order
            assert plans[s][2][d] == off
            off += int(plans[s][0][d])
        assert off == int(plans[d][3][0])                               # ... and fill exactly n_recv[d]
    assert len({int(p[4][0]) for p in plans}) == 1
    assert int(plans[0][4][0]) == max(int(p[3][0]) for p in plans)
    from thrill_b200 import capi
    assert capi.lib().tg_exchange_plan(0, 0, None, None, None, None, None, None) != 0
