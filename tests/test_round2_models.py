"""numpy models of the round-2 device-side arithmetic (documentation + brute-force checks of the formulas the kernels use; the
CUDA code itself is checked on the GPU against the oracle):
  * bit-granular prefix digits from the top varying bit (tg_radix_sort.cu prefix_sort_fast)
  * the interleaved tile list built on the device (tg_segmented.cuh seg_tiles_prepare/fill kernels) == the host-built list
  * splitters from p ordered sample lists by rank counting (tg_sample_sort.cu select_splitters_kernel) == FindAndSendSplitters
  * the splitter lookup table by the most significant key byte (SplitterDigit) == the plain (key, index) classification"""
import numpy as np
import pytest

from test_prefix_sort_model import HALO, prefix_digits_for


# ---- bit-granular prefix digits ------------------------------------------------------------------------------------------
def bitwise_prefix_sort_model(keys, top_bit):
    """K 8-bit digits directly below `top_bit`, most significant first then LSD inside (any stable order of the passes gives the
    same result: modelled as one stable sort by the prefix), finishing pass on the runs of equal prefixes"""
    n = len(keys)
    K = prefix_digits_for(n)
    if (top_bit + 7) // 8 < K + 2 or top_bit - 8 * K < 0:
        return np.sort(keys, kind="stable"), None
    diff = int(np.bitwise_or.reduce(keys) ^ np.bitwise_and.reduce(keys))
    tb_act = diff.bit_length()
    if tb_act > top_bit:
        return None, "bits above the assumed top digit vary"
    shift = np.uint64(top_bit - 8 * K)
    pre = keys >> shift
    order = np.argsort(pre, kind="stable")
    k, pre = keys[order], pre[order]
    starts = np.flatnonzero(np.concatenate(([True], pre[1:] != pre[:-1])))
    ends = np.concatenate((starts[1:], [n]))
    if (ends - starts).max() > HALO:
        return None, "run too long"
    out = k.copy()
    for s, e in zip(starts, ends):
        if e - s > 1:
            out[s:e] = np.sort(k[s:e], kind="stable")
    return out, None


@pytest.mark.parametrize("p,r", [(8, 3), (8, 0), (8, 7), (3, 1), (16, 9)])
def test_bitwise_prefix_digits_on_a_worker_key_range(p, r):
    """a worker of a p-GPU sort holds the keys between its two splitters: the top varying bit follows from them and is not byte
    aligned; the K digits below it separate the keys as well as the byte-aligned ones do on the whole key space"""
    rs = np.random.RandomState(17 * p + r)
    n = 200000
    lo, hi = (r << 64) // p, ((r + 1) << 64) // p - 1
    keys = (lo + (rs.randint(0, 1 << 62, size=n, dtype=np.int64).astype(object) * 4 % (hi - lo + 1))).astype(np.uint64)
    top_bit = (lo ^ hi).bit_length()
    out, why = bitwise_prefix_sort_model(keys, top_bit)
    assert out is not None, why
    assert np.array_equal(out, np.sort(keys))
    # a too-low hint is detected (the pass on the assumed top digit would not be the most significant one)
    if top_bit > 40:
        bad, why = bitwise_prefix_sort_model(keys, top_bit - 3)
        assert bad is None and "vary" in why


# ---- device-built interleaved tile list -----------------------------------------------------------------------------------
def host_tile_list(seg_size, tile):
    """build_tile_list (tg_segmented.cuh): round r holds the r-th tile of every segment that has one, segments ordered by tile
    count (descending, ties by index)"""
    nt = [(s + tile - 1) // tile for s in seg_size]
    row0 = np.concatenate(([0], np.cumsum(nt)[:-1]))
    start = np.concatenate(([0], np.cumsum(seg_size)[:-1]))
    order = sorted(range(len(seg_size)), key=lambda s: (-nt[s], s))
    out = []
    for r in range(max(nt) if nt else 0):
        for sg in order:
            if nt[sg] <= r:
                break
            off = r * tile
            out.append((start[sg] + off, min(tile, seg_size[sg] - off), row0[sg] + r, (sg << 20) | r))
    return out


def device_tile_list(seg_size, tile):
    """seg_tiles_prepare_kernel + seg_tiles_fill_kernel: position of tile (s, r) = A(r) + sortrank[s], A(r) = r * C(r) + (T - P[C(r)])"""
    S = len(seg_size)
    nt = np.array([(s + tile - 1) // tile for s in seg_size], dtype=np.int64)
    row0 = np.concatenate(([0], np.cumsum(nt)[:-1]))
    start = np.concatenate(([0], np.cumsum(seg_size)[:-1]))
    sortrank = np.array([sum(1 for q in range(S) if nt[q] > nt[s] or (nt[q] == nt[s] and q < s)) for s in range(S)])
    snt = np.zeros(S, dtype=np.int64)
    snt[sortrank] = nt
    P = np.concatenate(([0], np.cumsum(snt)))
    T = int(P[S])
    out = [None] * T
    for row in range(T):
        sg = int(np.searchsorted(row0, row, side="right") - 1)          # last segment with row0 <= row
        r = row - row0[sg]
        C = int(np.sum(snt > r))
        pos = r * C + (T - P[C]) + sortrank[sg]
        off = r * tile
        assert out[pos] is None
        out[pos] = (start[sg] + off, min(tile, seg_size[sg] - off), row, (sg << 20) | r)
    return out


@pytest.mark.parametrize("seed", range(6))
def test_device_tile_list_equals_the_host_list(seed):
    rs = np.random.RandomState(seed)
    S, tile = 256, 4096
    seg = rs.randint(0, 40 * tile, size=S)
    seg[rs.randint(0, S, size=40)] = 0                                  # empty buckets
    if seed % 2:
        seg[rs.randint(0, S)] = 700 * tile + 5                          # one dominant bucket (skewed digit)
    if seed == 5:
        seg[:] = 0; seg[17] = 3 * tile; seg[255] = 1
    assert device_tile_list(list(seg), tile) == host_tile_list(list(seg), tile)


# ---- splitters from ordered sample lists ----------------------------------------------------------------------------------
def reference_splitters(samples, p):
    """FindAndSendSplitters (api/sort.hpp:357-372): sort all (key, global index) samples, take samples[floor(i * S / p)]"""
    allv = sorted(samples)
    step = float(len(allv)) / float(p)
    return [allv[int(i * step)] for i in range(1, p)]


def device_splitters(lists, n_of, p):
    """select_splitters_kernel: sample (q, j) of the ordered list q has global rank j + sum over the other lists of the samples
    below it; local indices become global by the prefix of the shard sizes"""
    prefix = np.concatenate(([0], np.cumsum(n_of)))
    glob = [[(k, i + int(prefix[q])) for (k, i) in lst] for q, lst in enumerate(lists)]
    S = sum(len(l) for l in glob)
    want = {int(i * (float(S) / float(p))): i for i in range(p - 1, 0, -1)}      # (several i may want the same position)
    spl = [None] * (p - 1)
    for q, lst in enumerate(glob):
        for j, me in enumerate(lst):
            rank = j + sum(int(np.searchsorted(np.array([a * (1 << 70) + b for a, b in other], dtype=object),
                                               me[0] * (1 << 70) + me[1])) for w, other in enumerate(glob) if w != q)
            for i in range(1, p):
                if int(i * (float(S) / float(p))) == rank:
                    spl[i - 1] = me
    return spl


@pytest.mark.parametrize("p,dups", [(2, False), (5, False), (8, True), (3, True)])
def test_device_splitter_selection_equals_the_reference_rule(p, dups):
    rs = np.random.RandomState(100 * p + dups)
    n_of = rs.randint(0 if p > 2 else 50, 400, size=p)
    lists = []
    for q in range(p):
        ns = min(int(n_of[q]), 40)
        idx = rs.randint(0, max(int(n_of[q]), 1), size=ns)              # drawn with replacement: the same index may repeat
        keys = rs.randint(0, 6 if dups else 1 << 40, size=max(int(n_of[q]), 1))
        lists.append(sorted((int(keys[i]), int(i)) for i in idx))       # the sample leaves the worker ordered by (key, local index)
    prefix = np.concatenate(([0], np.cumsum(n_of)))
    flat = [(k, i + int(prefix[q])) for q, lst in enumerate(lists) for (k, i) in lst]
    if not flat:
        return
    assert device_splitters(lists, n_of, p) == reference_splitters(flat, p)


# ---- classification with the top-byte lookup table ------------------------------------------------------------------------
def classify_plain(key, gidx, spl):
    return sum(1 for s in spl if s < (key, gidx))                       # #splitters (key, idx) < (item key, global index)


def classify_lut(key, gidx, spl, lut_lo, lut_hi):
    tb = key >> 56
    lo, hi = lut_lo[tb], lut_hi[tb]
    while lo < hi:                                                      # only where the byte range holds a splitter
        mid = (lo + hi) // 2
        if spl[mid] < (key, gidx):
            lo = mid + 1
        else:
            hi = mid
    return lo


@pytest.mark.parametrize("case", ["uniform", "zipf_like", "ties"])
def test_top_byte_lookup_table_classification(case):
    rs = np.random.RandomState(len(case))
    p, n = 8, 20000
    if case == "uniform":
        keys = rs.randint(0, 1 << 63, size=n, dtype=np.int64).astype(np.uint64) * np.uint64(2)
    elif case == "zipf_like":
        keys = (rs.zipf(1.3, size=n) % (1 << 26)).astype(np.uint64)     # every key and every splitter share the top byte 0
    else:
        keys = rs.randint(0, 4, size=n).astype(np.uint64) << np.uint64(60)
    items = [(int(k), i) for i, k in enumerate(keys)]
    sample = sorted(items[i] for i in rs.randint(0, n, size=400))
    spl = [sample[int(i * len(sample) / p)] for i in range(1, p)]
    lut_lo = [sum(1 for s in spl if (s[0] >> 56) < b) for b in range(256)]
    lut_hi = [sum(1 for s in spl if (s[0] >> 56) <= b) for b in range(256)]
    for key, gidx in items[::7]:
        assert classify_lut(key, gidx, spl, lut_lo, lut_hi) == classify_plain(key, gidx, spl)
