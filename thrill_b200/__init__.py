"""thrill_b200 — B200-native Sort / ReduceByKey hot path for Thrill (drop-in for those two operators).

The product is the CUDA library thrill_b200/csrc/libthrill_gpu.so behind the C ABI of include/thrill_gpu.h.
This Python package is only the host-side mirror used by the tests and bench.py: `capi` binds the C ABI
with ctypes, `api` mirrors the reference's DIA<T> operator interface for the path (Sort, ReduceByKey,
ReducePair).  There is no CPU fallback: importing `capi` without the built library raises.
"""
__all__ = ["capi", "api"]
