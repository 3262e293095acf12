"""The C-ABI library loads, exports every symbol include/gpsat.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200 import binding
from gpupsat_b200.instances import random_ksat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpsat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpsat_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(binding.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []


def test_struct_sizes_match_header():
    assert ctypes.sizeof(binding.GpsatOpts) == binding.default_opts().struct_size
    assert binding.RECORD_DTYPE.itemsize == 80
    assert ctypes.sizeof(binding.GpsatStats) == 14 * 8 + 8 + 6 * 4 + 8 + 8 + 8 + 8


def test_no_cpu_fallback(request):
    from tests import conftest
    if conftest.HAS_GPU:
        pytest.skip("a GPU is present")
    offs, lits = random_ksat(20, 91, 0)
    with pytest.raises(g.GpsatError) as e:
        g.Solver(20, offs, lits)
    assert e.value.code == binding.E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_unpreprocessed_formulas():
    # unit clause / repeated variable must go through gpsat_cnf_preprocess first (two-watched literals need >= 2 lits)
    from gpupsat_b200.binding import lib, _p
    h = ctypes.c_void_p()
    offs = np.array([0, 1, 3], dtype=np.int64)
    lits = np.array([1, 2, 4], dtype=np.int32)
    rc = lib().gpsat_create(ctypes.byref(h), 3, 2, _p(offs), _p(lits), None)
    assert rc in (-1, -2)          # E_ARG with a GPU, E_NO_DEVICE without (device check comes first)


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "gpsat_ref" not in out and "emu" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_run" not in syms and "gpsat_emu_run" not in syms and "ref_solve" not in syms
