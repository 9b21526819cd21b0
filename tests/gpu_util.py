"""helpers shared by the -m gpu tests (all calls go through the C ABI via thrill_b200.capi)."""
import ctypes as C

import numpy as np


def make_blocks(capi, arr, block_bytes):
    """split a contiguous host array into tg_block views of `block_bytes` (items may straddle blocks)"""
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    n = len(raw)
    nb = max(1, (n + block_bytes - 1) // block_bytes) if n else 0
    blocks = (capi.Block * max(nb, 1))()
    for i in range(nb):
        lo = i * block_bytes
        hi = min(n, lo + block_bytes)
        blocks[i].data = raw.ctypes.data + lo
        blocks[i].bytes = hi - lo
    return blocks, nb, raw


def u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))
