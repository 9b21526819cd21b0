"""ctypes binding of oracle/_build/libthrill_oracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's Sort / ReduceByKey hot path
(oracle/thrill_oracle.c).  It is the checker for the CUDA path; the product package
(thrill_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libthrill_oracle.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "thrill_ref_driver")

KEY_UINT_LE, KEY_BYTES_BE = 0, 1
OP_SUM_F64, OP_SUM_U64, OP_MIN_U64, OP_MAX_U64, OP_MIN_F64, OP_MAX_F64, OP_FIRST = range(7)

KV = np.dtype([("key", "<u8"), ("val", "<u8")])
BLOCK_META = np.dtype([("begin", "<u8"), ("end", "<u8"), ("first_item", "<u8"), ("num_items", "<u8")])


class KVStruct(C.Structure):
    _fields_ = [("key", C.c_uint64), ("val", C.c_uint64)]


class KeyDesc(C.Structure):
    _fields_ = [("item_bytes", C.c_uint32), ("key_offset", C.c_uint32),
                ("key_bytes", C.c_uint32), ("key_kind", C.c_uint32)]


U64_DESC = KeyDesc(8, 0, 8, KEY_UINT_LE)
KV_DESC = KeyDesc(16, 0, 8, KEY_UINT_LE)
RECORD_DESC = KeyDesc(100, 0, 10, KEY_BYTES_BE)

_lib = None


def build():
    src = os.path.join(ORACLE_DIR, "thrill_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        u64, dbl, vp, i32, u32 = C.c_uint64, C.c_double, C.c_void_p, C.c_int, C.c_uint32
        L.to_splitmix64.restype = u64; L.to_splitmix64.argtypes = [u64]
        L.to_hash128to64.restype = u64; L.to_hash128to64.argtypes = [u64, u64]
        L.to_sample_size.restype = u64; L.to_sample_size.argtypes = [u64, dbl]
        L.to_zipf_rank.restype = u64; L.to_zipf_rank.argtypes = [vp, u64, dbl]
        L.to_select_splitters.restype = u64
        L.to_select_splitters.argtypes = [vp, u64, u64, C.POINTER(KeyDesc), vp]
        L.to_reduce_pre_phase.restype = u64
        L.to_reduce_pre_phase.argtypes = [vp, u64, u64, u64, i32, vp, vp, u64]
        L.to_reduce_post_phase.restype = u64
        L.to_reduce_post_phase.argtypes = [vp, u64, u64, i32, vp, u64, vp]
        L.to_reduce_operator.restype = u64
        L.to_reduce_operator.argtypes = [vp, vp, u32, u64, i32, vp, vp]
        L.to_reduce_simple.restype = u64; L.to_reduce_simple.argtypes = [vp, u64, i32, vp]
        L.to_reduce_to_index.restype = u64; L.to_reduce_to_index.argtypes = [vp, u64, u64, KVStruct, i32, vp]
        L.to_file_layout.restype = u64; L.to_file_layout.argtypes = [u64, u32, u64, u64, vp, u64]
        L.to_gen_sort_uniform.argtypes = [vp, u64, u64, u64]
        L.to_gen_reduce_uniform.argtypes = [vp, u64, u64, u64, u64, i32]
        L.to_zipf_build_cdf.argtypes = [vp, u64, dbl]
        L.to_gen_sort_zipf.argtypes = [vp, u64, u64, u64, vp, u64]
        L.to_gen_reduce_zipf.argtypes = [vp, u64, u64, u64, vp, u64, i32]
        L.to_gen_records.argtypes = [vp, u64, u64, u64]
        L.to_reduce_by_hash.argtypes = [u64, u64, u64, vp, vp]
        L.to_hash_partition_ids.argtypes = [vp, u64, u64, u64, u64, vp]
        L.to_sort_items.argtypes = [vp, u64, C.POINTER(KeyDesc)]
        L.to_build_tree.argtypes = [vp, u64, C.POINTER(KeyDesc), vp]
        L.to_classify.argtypes = [vp, u64, u64, vp, u64, u64, vp, C.POINTER(KeyDesc), vp]
        L.to_multiway_merge.argtypes = [vp, vp, u32, C.POINTER(KeyDesc), i32, vp]
        L.to_sort_operator.argtypes = [vp, vp, u32, C.POINTER(KeyDesc), i32, u64, vp, vp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- generators -------------------------------------------------------------
def splitmix64(x):
    return lib().to_splitmix64(C.c_uint64(x & (2**64 - 1)))


def gen_sort_uniform(begin, n, seed=42):
    out = np.empty(n, dtype=np.uint64)
    lib().to_gen_sort_uniform(_p(out), begin, n, seed)
    return out


def gen_reduce_uniform(begin, n, seed=42, universe=1 << 26, exact=0):
    out = np.empty(n, dtype=KV)
    lib().to_gen_reduce_uniform(_p(out), begin, n, seed, universe, exact)
    return out


def zipf_cdf(universe, s=1.0):
    cdf = np.empty(universe, dtype=np.float64)
    lib().to_zipf_build_cdf(_p(cdf), universe, s)
    return cdf


def gen_sort_zipf(begin, n, cdf, seed=42):
    out = np.empty(n, dtype=np.uint64)
    lib().to_gen_sort_zipf(_p(out), begin, n, seed, _p(cdf), len(cdf))
    return out


def gen_reduce_zipf(begin, n, cdf, seed=42, exact=0):
    out = np.empty(n, dtype=KV)
    lib().to_gen_reduce_zipf(_p(out), begin, n, seed, _p(cdf), len(cdf), exact)
    return out


def gen_records(begin, n, seed=42):
    out = np.empty((n, 100), dtype=np.uint8)
    lib().to_gen_records(_p(out), begin, n, seed)
    return out


# ---- hashing ---------------------------------------------------------------
def hash128to64(upper, lower):
    return lib().to_hash128to64(upper, lower)


def hash_partition_ids(keys, num_partitions, salt=0):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.empty(len(keys), dtype=np.uint32)
    lib().to_hash_partition_ids(_p(keys), len(keys), 8, salt, num_partitions, _p(out))
    return out


# ---- sort ------------------------------------------------------------------
def sample_size(count, imbalance=0.1):
    return lib().to_sample_size(count, imbalance)


def _as_items(items, desc):
    a = np.ascontiguousarray(items)
    n = a.nbytes // desc.item_bytes
    return a, n


def sort_items(items, desc=U64_DESC):
    a = np.array(items, copy=True)
    a, n = _as_items(a, desc)
    lib().to_sort_items(_p(a), n, C.byref(desc))
    return a


def pack_samples(items, gidx, desc=U64_DESC):
    """(item, u64 global index) packed back to back, as SampleIndexPair on the wire."""
    items = np.ascontiguousarray(items).view(np.uint8).reshape(len(gidx), desc.item_bytes)
    out = np.empty((len(gidx), desc.item_bytes + 8), dtype=np.uint8)
    out[:, :desc.item_bytes] = items
    out[:, desc.item_bytes:] = np.ascontiguousarray(gidx, dtype=np.uint64).view(np.uint8).reshape(-1, 8)
    return out


def select_splitters(samples_packed, p, desc=U64_DESC):
    s = np.array(samples_packed, copy=True)
    out = np.zeros((max(p - 1, 0) + 1, desc.item_bytes + 8), dtype=np.uint8)
    n = lib().to_select_splitters(_p(s), len(s), p, C.byref(desc), _p(out))
    return out[:n]


def pad_splitters(splitters, p, desc=U64_DESC):
    k = 1
    while k < p:
        k *= 2
    spl = [splitters[i] for i in range(len(splitters))]
    while len(spl) < k - 1:
        spl.append(spl[-1])
    arr = np.zeros((k, desc.item_bytes + 8), dtype=np.uint8)
    if spl:
        arr[:len(spl)] = np.stack(spl)
    return arr, k


def build_tree(splitters_padded, k, desc=U64_DESC):
    tree = np.zeros((k + 1, desc.item_bytes), dtype=np.uint8)
    lib().to_build_tree(_p(np.ascontiguousarray(splitters_padded)), k, C.byref(desc), _p(tree))
    return tree


def classify(items, prefix_items, tree, k, splitters_padded, desc=U64_DESC):
    a, n = _as_items(items, desc)
    log_k = k.bit_length() - 1
    out = np.empty(n, dtype=np.uint32)
    lib().to_classify(_p(a), n, prefix_items, _p(np.ascontiguousarray(tree)), k, log_k,
                      _p(np.ascontiguousarray(splitters_padded)), C.byref(desc), _p(out))
    return out


def multiway_merge(runs, desc=U64_DESC, stable=False):
    runs = [np.ascontiguousarray(r) for r in runs]
    k = len(runs)
    ptrs = (C.c_void_p * k)(*[r.ctypes.data for r in runs])
    counts = np.array([r.nbytes // desc.item_bytes for r in runs], dtype=np.uint64)
    total = int(counts.sum())
    out = np.empty(total * desc.item_bytes, dtype=np.uint8)
    lib().to_multiway_merge(ptrs, _p(counts), k, C.byref(desc), int(stable), _p(out))
    return out


def sort_operator(items, local_counts, desc=U64_DESC, stable=False, rng_seed=1):
    a, n = _as_items(items, desc)
    lc = np.asarray(local_counts, dtype=np.uint64)
    assert int(lc.sum()) == n
    out = np.empty(n * desc.item_bytes, dtype=np.uint8)
    oc = np.zeros(len(lc), dtype=np.uint64)
    lib().to_sort_operator(_p(a), _p(lc), len(lc), C.byref(desc), int(stable), rng_seed, _p(out), _p(oc))
    return out, oc


# ---- reduce ----------------------------------------------------------------
def reduce_pre_phase(kv, p, limit_memory_bytes, op):
    kv = np.ascontiguousarray(kv, dtype=KV)
    out = np.empty(len(kv) + 1, dtype=KV)
    part = np.empty(len(kv) + 1, dtype=np.uint32)
    n = lib().to_reduce_pre_phase(_p(kv), len(kv), p, limit_memory_bytes, op, _p(out), _p(part), len(out))
    return out[:n], part[:n]


def reduce_post_phase(kv, limit_memory_bytes, op):
    kv = np.ascontiguousarray(kv, dtype=KV)
    out = np.empty(len(kv) + 1, dtype=KV)
    iters = C.c_uint64(0)
    n = lib().to_reduce_post_phase(_p(kv), len(kv), limit_memory_bytes, op, _p(out), len(out), C.byref(iters))
    return out[:n], iters.value


def reduce_operator(kv, local_counts, mem_limit_bytes, op):
    kv = np.ascontiguousarray(kv, dtype=KV)
    lc = np.asarray(local_counts, dtype=np.uint64)
    out = np.empty(len(kv) + 1, dtype=KV)
    oc = np.zeros(len(lc), dtype=np.uint64)
    n = lib().to_reduce_operator(_p(kv), _p(lc), len(lc), mem_limit_bytes, op, _p(out), _p(oc))
    return out[:n], oc


def reduce_simple(kv, op):
    kv = np.ascontiguousarray(kv, dtype=KV)
    out = np.empty(max(len(kv), 1), dtype=KV)
    n = lib().to_reduce_simple(_p(kv), len(kv), op, _p(out))
    return out[:n]


def reduce_to_index(kv, size, op, neutral=(0, 0)):
    """dense ReduceToIndex result (api/reduce_to_index.hpp): `size` items, neutral where no item has that index"""
    kv = np.ascontiguousarray(kv, dtype=KV)
    out = np.empty(max(size, 1), dtype=KV)
    bad = lib().to_reduce_to_index(_p(kv), len(kv), size, KVStruct(int(neutral[0]), int(neutral[1])), op, _p(out))
    assert bad == 0, "%d indices out of range" % bad
    return out[:size]


# ---- layout ----------------------------------------------------------------
def file_layout(num_items, item_bytes, start_block_size=4096, max_block_size=2 << 20):
    cap = 64 + (num_items * item_bytes) // max(1, min(start_block_size, max_block_size))
    out = np.zeros(cap, dtype=BLOCK_META)
    n = lib().to_file_layout(num_items, item_bytes, start_block_size, max_block_size, _p(out), cap)
    assert n <= cap
    return out[:n]


# ---- the real reference (only where the prebuilt binary exists) -------------
def have_ref_driver():
    return os.path.exists(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


def run_ref_driver(workers=2, **kw):
    env = dict(os.environ, THRILL_NET="mock", THRILL_LOCAL="1", THRILL_WORKERS_PER_HOST=str(workers), THRILL_LOG="")
    args = [REF_DRIVER] + ["%s=%s" % (k, v) for k, v in kw.items()]
    res = subprocess.run(args, env=env, capture_output=True, text=True, timeout=1800)
    if res.returncode != 0:
        raise RuntimeError("thrill_ref_driver failed: %s\n%s" % (res.returncode, res.stderr[-2000:]))
    times = [float(l.rsplit("time=", 1)[1]) for l in res.stdout.splitlines() if l.startswith("RESULT")]
    return times, res.stdout
