"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the CPU oracle (oracle/_ref/libgpsat_oracle.so) and of the
reference's own solver classes built for the host (oracle/_ref/libgpsat_ref.so, libgpsat_ref_nocap.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

SAT, UNSAT, UNDEF = 0, 1, 2

RECORD_DTYPE = np.dtype([
    ("status", np.int32), ("reserved", np.int32), ("decisions", np.int64), ("implications", np.int64),
    ("conflicts", np.int64), ("learnt_clauses", np.int64), ("learnt_literals", np.int64), ("restarts", np.int64),
    ("watchers_visited", np.int64), ("clause_words_read", np.int64), ("learnt_hash", np.int64)])


class Quiet:
    """Silences the C-level stdout of the reference host build (its library code printf()s)."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *a):
        import sys
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def build(force: bool = False) -> None:
    """Compile the oracle (always possible) and the reference host build (only where /root/reference exists)."""
    so = os.path.join(REF_DIR, "libgpsat_oracle.so")
    src = os.path.join(HERE, "gpsat_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "_ref/libgpsat_oracle.so"], stdout=subprocess.DEVNULL)
    ref_src = os.environ.get("GPSAT_REFERENCE_SRC", "/root/reference/src")
    if os.path.isdir(ref_src) and (force or not os.path.exists(os.path.join(REF_DIR, "libgpsat_ref.so"))):
        subprocess.check_call([os.path.join(HERE, "ref", "build_ref.sh")], stdout=subprocess.DEVNULL)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The restatement (oracle/gpsat_oracle.cpp)."""

    def __init__(self, n_vars, offsets, lits):
        build()
        self.lib = C.CDLL(os.path.join(REF_DIR, "libgpsat_oracle.so"))
        self.lib.oracle_open.restype = C.c_void_p
        self.lib.oracle_open.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.oracle_close.argtypes = [C.c_void_p]
        self.lib.oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.oracle_eval_clauses.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lits = np.ascontiguousarray(lits, dtype=np.int32)
        self.n_vars = int(n_vars)
        self.n_clauses = len(self.offsets) - 1
        self.h = C.c_void_p(self.lib.oracle_open(self.n_vars, self.n_clauses, _p(self.offsets), _p(self.lits)))

    def close(self):
        if self.h:
            self.lib.oracle_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, cube_offsets, cube_lits, *, mode=0, decision=1, restart_first=100, restart_factor=1.3,
            max_iterations=0, max_conflicts=0, max_learnts_first=None, learnt_refs_cap=16384,
            arena_words=1 << 19, stop_on_sat=True, implied_stride=None):
        co = np.ascontiguousarray(cube_offsets, dtype=np.int64)
        cl = np.ascontiguousarray(cube_lits, dtype=np.int32)
        n_cubes = len(co) - 1
        if max_learnts_first is None:
            max_learnts_first = default_max_learnts(self.n_clauses, learnt_refs_cap, self.n_vars)
        ip = np.array([mode, decision, restart_first, max_iterations, max_learnts_first, learnt_refs_cap], dtype=np.int32)
        rec = np.zeros(n_cubes, dtype=RECORD_DTYPE)
        model = np.zeros(max(self.n_vars, 1), dtype=np.uint8)
        sat_job = C.c_int32(-1)
        stride = self.n_vars if implied_stride is None else implied_stride
        implied = n_implied = confl = None
        if mode != 0:
            implied = np.full(max(n_cubes * stride, 1), -1, dtype=np.int32)
            n_implied = np.zeros(n_cubes, dtype=np.int32)
            confl = np.full(n_cubes, -1, dtype=np.int64)
        self.lib.oracle_run(self.h, _p(ip), C.c_float(restart_factor), max_conflicts, arena_words, n_cubes, _p(co),
                            _p(cl), _p(rec), _p(model), C.byref(sat_job), 1 if stop_on_sat else 0, _p(implied), stride,
                            _p(n_implied), _p(confl))
        out = {"records": rec, "sat_job": sat_job.value, "model": model[: self.n_vars]}
        if mode != 0:
            out.update(implied=implied.reshape(n_cubes, stride) if stride else implied, n_implied=n_implied,
                       conflict_clause=confl)
        return out

    def eval_clauses(self, assignment):
        a = np.ascontiguousarray(assignment, dtype=np.uint8).reshape(-1, self.n_vars)
        status = np.zeros((a.shape[0], self.n_clauses), dtype=np.int32)
        unit = np.zeros((a.shape[0], self.n_clauses), dtype=np.int32)
        self.lib.oracle_eval_clauses(self.h, a.shape[0], _p(a), _p(status), _p(unit))
        return status, unit


def default_max_learnts(n_clauses, refs_cap, n_vars):
    """The product's default for gpsat_solve_params.max_learnts_first (gpsat_api.cu: default_max_learnts)."""
    v = max(n_clauses // 3, 300)
    return max(min(v, refs_cap - n_vars - 2), 1)


class Reference:
    """The reference's own classes compiled for the host (oracle/ref/ref_driver.cpp)."""

    def __init__(self, offsets, lits, capacity=None, nocap=False):
        build()
        name = "libgpsat_ref_nocap.so" if nocap else "libgpsat_ref.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        for f in ("ref_n_vars", "ref_status_after_preprocessing", "ref_n_clauses", "ref_n_solved_literals"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_n_literals.argtypes = [C.c_void_p]
        L.ref_n_literals.restype = C.c_int64
        L.ref_get_solved_literals.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_formula.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_cubes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.ref_cubes_simple.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ref_clause_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_counter_implications.restype = C.c_longlong
        L.ref_counter_decisions.restype = C.c_longlong
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lits = np.ascontiguousarray(lits, dtype=np.int32)
        m = len(offsets) - 1
        self.h = C.c_void_p(L.ref_open(m, _p(offsets), _p(lits), capacity if capacity is not None else m + 1))
        self.n_vars = L.ref_n_vars(self.h)
        self.status = L.ref_status_after_preprocessing(self.h)

    def formula(self):
        m = self.lib.ref_n_clauses(self.h)
        nl = self.lib.ref_n_literals(self.h)
        off = np.zeros(m + 1, dtype=np.int64)
        lits = np.zeros(max(nl, 1), dtype=np.int32)
        self.lib.ref_get_formula(self.h, _p(off), _p(lits))
        return off, lits[:nl]

    def solved_literals(self):
        n = self.lib.ref_n_solved_literals(self.h)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.ref_get_solved_literals(self.h, _p(out))
        return out[:n]

    def cubes(self, blocks, threads, strategy=0):
        k = C.c_int(0)
        n = self.lib.ref_cubes(self.h, blocks, threads, strategy, C.byref(k), None, 0)
        out = np.zeros(max(n * k.value, 1), dtype=np.int32)
        self.lib.ref_cubes(self.h, blocks, threads, strategy, C.byref(k), _p(out), n * k.value)
        return out[: n * k.value].reshape(n, k.value)

    def cubes_simple(self):
        """SimpleJobChooser (JobsManager/SimpleJobChooser.cu:22-75): the generator behind USE_SIMPLE_JOBS_GENERATION."""
        k = C.c_int(0)
        n = self.lib.ref_cubes_simple(self.h, C.byref(k), None, 0)
        out = np.zeros(max(n * k.value, 1), dtype=np.int32)
        self.lib.ref_cubes_simple(self.h, C.byref(k), _p(out), n * k.value)
        return out[: n * k.value].reshape(n, k.value)

    def clause_status(self, assignment):
        """VariablesStateHandler::clause_status of every clause under `assignment` (0 true, 1 false, 2 unassigned)."""
        a = np.ascontiguousarray(assignment, dtype=np.uint8)
        m = self.lib.ref_n_clauses(self.h)
        status = np.zeros(m, dtype=np.int32)
        unit = np.zeros(m, dtype=np.int32)
        got = self.lib.ref_clause_status(self.h, _p(a), _p(status), _p(unit))
        assert got == m
        return status, unit

    def propagate(self, cube):
        cube = np.ascontiguousarray(cube, dtype=np.int32)
        st = C.c_int32(-1)
        n = C.c_int32(0)
        buf = np.zeros(self.n_vars + 1, dtype=np.int32)
        rc = self.lib.ref_propagate(self.h, _p(cube), len(cube), C.byref(st), _p(buf), C.byref(n))
        assert rc == 0
        return st.value, buf[: n.value].copy()

    def solve(self, cube=()):
        cube = np.ascontiguousarray(cube, dtype=np.int32)
        st = C.c_int32(-1)
        n = C.c_int32(0)
        buf = np.zeros(self.n_vars + 1, dtype=np.int32)
        rc = self.lib.ref_solve(self.h, _p(cube) if len(cube) else None, len(cube), C.byref(st), _p(buf), C.byref(n))
        assert rc == 0
        return st.value, buf[: n.value].copy()

    def counters(self):
        return self.lib.ref_counter_implications(), self.lib.ref_counter_decisions()

    def reset_counters(self):
        self.lib.ref_counters_reset()
