"""CPU-side checks of the C ABI: the library loads, exports every symbol include/thrill_gpu.h declares,
fails loudly without a GPU (no fallback), and its pure-host functions agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from thrill_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "thrill_gpu.h")).read()
    declared = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tg_ctx"}
    L = capi.lib()
    bound = {name for name, _, _ in capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(L, name)
    assert L.tg_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.ThrillGpuError) as e:
        capi.Ctx(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "thrill_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_lib|thrill_oracle|libthrill_oracle|oracle/_", src):
                    if not (f == "capi.py" or f == "__init__.py"):
                        bad.append(f)
                    elif re.search(r"import\s+oracle|CDLL\([^)]*oracle", src):
                        bad.append(f)
    assert not bad, bad


def test_sample_size_matches_oracle():
    L = capi.lib()
    for n in [1, 2, 3, 100, 4097, 10**6, 10**8, 125000000, 2**30 - 1]:
        assert L.tg_sample_size(n) == O.sample_size(n)


@pytest.mark.parametrize("p", [1, 2, 3, 5, 8, 16])
def test_select_splitters_matches_oracle(p):
    rng = np.random.RandomState(p)
    n = 5000
    keys = rng.randint(0, 40, size=n).astype(np.uint64)            # many ties: index order matters
    idx = rng.permutation(10**6)[:n].astype(np.uint64)
    samples = O.pack_samples(keys, idx)
    want = O.select_splitters(samples, p)
    mine = np.array(samples, copy=True)
    out = np.zeros((max(p - 1, 1), 16), dtype=np.uint8)
    d = capi.u64_desc()
    st = capi.lib().tg_select_splitters(C.byref(d), mine.ctypes.data, n, p, out.ctypes.data)
    assert st == 0
    assert np.array_equal(out[:p - 1], want)


def test_select_splitters_record_keys():
    rng = np.random.RandomState(9)
    n, p = 3000, 7
    rec = rng.randint(0, 256, size=(n, 16)).astype(np.uint8)
    rec[:, 0] = rng.randint(0, 2, size=n)
    idx = np.arange(n, dtype=np.uint64) * 3
    od = O.KeyDesc(16, 0, 10, O.KEY_BYTES_BE)
    samples = O.pack_samples(rec, idx, od)
    want = O.select_splitters(samples, p, od)
    mine = np.array(samples, copy=True)
    out = np.zeros((p - 1, 24), dtype=np.uint8)
    d = capi.KeyDesc(16, 0, 10, capi.KEY_BYTES_BE, 0, 0)
    assert capi.lib().tg_select_splitters(C.byref(d), mine.ctypes.data, n, p, out.ctypes.data) == 0
    assert np.array_equal(out, want)


@pytest.mark.parametrize("n,ib,start,mx", [(47, 2, 4096, 16), (1000, 100, 4096, 2 << 20), (1 << 20, 8, 4096, 2 << 20),
                                           (0, 8, 4096, 2 << 20), (1, 100, 4096, 2 << 20), (12345, 16, 64, 1024)])
def test_file_geometry_matches_blockwriter_restatement(n, ib, start, mx):
    want = O.file_layout(n, ib, start, mx)
    cap = len(want) + 4
    geo = (capi.BlockGeom * cap)()
    nb = capi.lib().tg_file_geometry(n, ib, start, mx, geo, cap)
    assert nb == len(want)
    for i in range(nb):
        assert geo[i].bytes == int(want[i]["end"])
        assert geo[i].num_items == int(want[i]["num_items"])
        assert geo[i].first_item == int(want[i]["first_item"])


def test_header_is_plain_c(tmp_path):
    """include/thrill_gpu.h is the drop-in boundary: it must compile as C99 (no C++ types in the signatures) and link against
    the shared library from a C program"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "thrill_gpu.h"\nint main(void) { tg_dev_file f; tg_key_desc d; (void)f; (void)d; return tg_version() >= 100 ? 0 : 1; }\n')
    exe = str(tmp_path / "hdr")
    libdir = os.path.join(ROOT, "thrill_b200", "csrc")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                          "-L", libdir, "-lthrill_gpu", "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert subprocess.run([exe]).returncode == 0
