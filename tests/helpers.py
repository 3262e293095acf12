import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def golden():
    with open(os.path.join(HERE, "golden", "reference_outputs.json")) as f:
        return json.load(f)


def have_ref():
    from oracle import binding
    return os.path.exists(os.path.join(binding.REF_DIR, "libgpsat_ref.so"))


def cube_csr(cubes):
    n, k = cubes.shape
    return np.arange(0, n * k + 1, k, dtype=np.int64), np.ascontiguousarray(cubes.reshape(-1), dtype=np.int32)


def cube_checksum(cb):
    return int((cb.astype(np.int64) * (np.arange(cb.size).reshape(cb.shape) % 1009 + 1)).sum())


def model_from_lits(n_vars, lits):
    m = np.ones(n_vars, dtype=np.uint8)
    for x in lits:
        m[x >> 1] = x & 1
    return m
