#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/thrill_ref_driver, built
from /root/reference by oracle/ref/Makefile) on the deterministic inputs of SURVEY.md §8(d).

Run in the build container only (needs /root/reference for the build):  python tests/golden/make_golden.py
The fixtures pin oracle/thrill_oracle.c (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REDUCE_OUT = np.dtype([("key", "<u8"), ("val", "<u8"), ("worker", "<u8")])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert O.have_ref_driver(), "build oracle/_ref first: make -C oracle ref"
    tmp = tempfile.mkdtemp()
    out = os.path.join(tmp, "o.bin")
    g = {}
    # Sort, uniform u64 (cfg2 shape), full output small + digest large
    for name, n, w in (("sort_uniform_4096_w3", 4096, 3), ("sort_uniform_1000000_w4", 1000000, 4)):
        O.run_ref_driver(workers=w, op="sort_u64", n=n, out=out)
        a = np.fromfile(out, dtype=np.uint64)
        assert len(a) == n
        g[name + "_sha256"] = np.array(sha(a))
        if n <= 4096:
            g[name] = a
    # Sort, Zipf keys (massive duplicates: splitter tie-break path)
    O.run_ref_driver(workers=5, op="sort_u64", gen="zipf", universe=1024, n=50000, out=out)
    a = np.fromfile(out, dtype=np.uint64)
    g["sort_zipf_u1024_50000_w5_sha256"] = np.array(sha(a))
    g["sort_zipf_u1024_50000_w5_head"] = a[:64].copy()
    # ReducePair<u64,double>(plus), Zipf s=1.0 (cfg3 shape, reduced): sorted by key
    O.run_ref_driver(workers=3, op="reduce_f64", gen="zipf", universe=4096, n=200000, out=out)
    r = np.sort(np.fromfile(out, dtype=REDUCE_OUT), order="key")
    g["reduce_f64_zipf_u4096_200000_w3"] = r
    # exact mode (integer-valued doubles): bit-exact sums
    O.run_ref_driver(workers=4, op="reduce_f64", gen="zipf", universe=4096, n=200000, exact=1, out=out)
    g["reduce_f64_exact_zipf_u4096_200000_w4"] = np.sort(np.fromfile(out, dtype=REDUCE_OUT), order="key")
    # uniform keys, u64 sum
    O.run_ref_driver(workers=2, op="reduce_u64", gen="uniform", universe=3000, n=100000, out=out)
    g["reduce_u64_uniform_u3000_100000_w2"] = np.sort(np.fromfile(out, dtype=REDUCE_OUT), order="key")
    # ReduceToIndex (the PageRank step): dense result, exact-mode doubles (bit-exact) and real doubles (tolerance)
    for exact, w in ((1, 3), (0, 4)):
        O.run_ref_driver(workers=w, op="reduce_to_index", gen="zipf", universe=1000, n=20000, exact=exact, out=out)
        g["reduce_to_index_zipf_u1000_20000_exact%d_w%d" % (exact, w)] = np.fromfile(out, dtype=O.KV)
    # sparse: most indices have no item and keep the neutral element (0, 0.0)
    O.run_ref_driver(workers=5, op="reduce_to_index", gen="uniform", universe=50000, n=3000, exact=1, out=out)
    g["reduce_to_index_uniform_u50000_3000_exact1_w5"] = np.fromfile(out, dtype=O.KV)
    # TeraSort records
    O.run_ref_driver(workers=3, op="terasort", n=20000, out=out)
    t = np.fromfile(out, dtype=np.uint8).reshape(-1, 100)
    g["terasort_20000_w3_sha256"] = np.array(sha(t))
    g["terasort_20000_w3_head"] = t[:8].copy()
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **g)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"), {k: getattr(v, "shape", None) for k, v in g.items()})


if __name__ == "__main__":
    main()
