"""Shared helpers for the golden fixtures (tests/golden/reference_outputs.npz, made by make_golden.py
from outputs of the unmodified reference)."""
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REDUCE_OUT = np.dtype([("key", "<u8"), ("val", "<u8"), ("worker", "<u8")])


def golden():
    return np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


def golden_r2():
    """larger cases (digests only), made by tests/golden/make_golden_r2.py from the unmodified reference"""
    return np.load(os.path.join(HERE, "golden", "reference_outputs_r2.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
