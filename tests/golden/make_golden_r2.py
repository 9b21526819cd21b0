#!/usr/bin/env python
"""Round-2 fixtures (tests/golden/reference_outputs_r2.npz): larger cases of the UNMODIFIED reference
(oracle/_ref/thrill_ref_driver), digests only.  Run in the build container:  python tests/golden/make_golden_r2.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from golden_util import REDUCE_OUT, sha  # noqa: E402


def main():
    assert O.have_ref_driver(), "build oracle/_ref first: make -C oracle ref"
    out = os.path.join(tempfile.mkdtemp(), "o.bin")
    g = {}
    # TeraSort at 1e6 records (the output does not depend on the worker count)
    O.run_ref_driver(workers=4, op="terasort", n=1000000, out=out)
    t = np.fromfile(out, dtype=np.uint8).reshape(-1, 100)
    assert len(t) == 1000000
    g["terasort_1000000_sha256"] = np.array(sha(t))
    g["terasort_1000000_head"] = t[:4].copy()
    g["terasort_1000000_tail"] = t[-4:].copy()
    # Sort of 1e7 uniform keys
    O.run_ref_driver(workers=4, op="sort_u64", n=10000000, out=out)
    a = np.fromfile(out, dtype=np.uint64)
    g["sort_uniform_10000000_sha256"] = np.array(sha(a))
    # Sort of 5e6 Zipf(s=1, U=2^20) keys: heavy duplicates at scale (splitter tie-break on every worker boundary)
    O.run_ref_driver(workers=5, op="sort_u64", gen="zipf", universe=1 << 20, n=5000000, out=out)
    a = np.fromfile(out, dtype=np.uint64)
    g["sort_zipf_u2^20_5000000_sha256"] = np.array(sha(a))
    # ReducePair, Zipf(s=1, U=2^20), 4e6 records, exact mode: digest of the key-sorted (key, sum) pairs
    O.run_ref_driver(workers=4, op="reduce_f64", gen="zipf", universe=1 << 20, n=4000000, exact=1, out=out)
    r = np.sort(np.fromfile(out, dtype=REDUCE_OUT), order="key")
    kv = np.zeros(len(r), dtype=O.KV)
    kv["key"], kv["val"] = r["key"], r["val"]
    g["reduce_f64_exact_zipf_u2^20_4000000_sha256"] = np.array(sha(kv))
    g["reduce_f64_exact_zipf_u2^20_4000000_distinct"] = np.array(len(kv))
    np.savez_compressed(os.path.join(HERE, "reference_outputs_r2.npz"), **g)
    print("wrote reference_outputs_r2.npz", {k: (str(v) if v.ndim == 0 else v.shape) for k, v in g.items()})


if __name__ == "__main__":
    main()
