"""CPU model of the prefix sort (thrill_b200/csrc/tg_radix_sort.cu): ordering by the K most significant non-constant key
bytes with stable passes and finishing every run of equal prefixes by (full key, position) is the stable sort by the full
key; runs longer than the finishing pass's halo (64) must be detected.  numpy only — documents and checks the algorithm
the CUDA path implements (the CUDA path itself is checked on the GPU against the oracle, tests/test_gpu_radix.py)."""
import numpy as np
import pytest

HALO = 64


def prefix_digits_for(n):
    bits = 4
    while bits < 64 and (1 << bits) < n * 16:
        bits += 1
    return (bits + 7) // 8


def prefix_sort_model(keys):
    """returns (sorted keys or None if a run is too long, K, active byte positions)"""
    n = len(keys)
    b = keys.view(np.uint8).reshape(n, 8)
    active = [p for p in range(8) if b[:, p].min() != b[:, p].max()]
    K = prefix_digits_for(n)
    if len(active) < K + 2:
        return np.sort(keys, kind="stable"), K, active          # plain LSD over the active bytes
    order = np.arange(n)
    for p in active[-K:]:                                       # stable passes, least significant of the K first
        order = order[np.argsort(b[order, p], kind="stable")]
    k = keys[order]
    shift = np.uint64(8 * active[-K])
    pre = k >> shift
    starts = np.flatnonzero(np.concatenate(([True], pre[1:] != pre[:-1])))
    ends = np.concatenate((starts[1:], [n]))
    if (ends - starts).max() > HALO:
        return None, K, active
    out = k.copy()
    for s, e in zip(starts, ends):
        if e - s > 1:
            out[s:e] = np.sort(k[s:e], kind="stable")           # rank by (full key, position)
    return out, K, active


@pytest.mark.parametrize("n", [1000, 40000, 300000])
def test_model_uniform_keys_take_the_prefix_path(n):
    keys = np.random.RandomState(n).randint(0, 2**63 - 1, size=n, dtype=np.int64).astype(np.uint64)
    out, K, active = prefix_sort_model(keys)
    assert len(active) >= K + 2 and out is not None
    assert np.array_equal(out, np.sort(keys))


def test_model_small_keys_are_plain_lsd_and_long_runs_are_detected():
    small = np.random.RandomState(1).randint(0, 1 << 20, size=50000).astype(np.uint64)
    out, K, active = prefix_sort_model(small)
    assert len(active) < K + 2 and np.array_equal(out, np.sort(small))
    pool = np.random.RandomState(2).randint(0, 2**63 - 1, size=200, dtype=np.int64).astype(np.uint64)
    dup = np.repeat(pool, 150)
    np.random.RandomState(3).shuffle(dup)
    out, _, _ = prefix_sort_model(dup)
    assert out is None                                           # the CUDA path falls back to the plain LSD passes here


def test_model_sample_size_rule():
    assert [prefix_digits_for(n) for n in (1 << 15, 10**6, 10**8, 10**9)] == [3, 3, 4, 5]
