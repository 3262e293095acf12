"""ctypes binding of libgpsat.so (include/gpsat.h).  Plumbing only: every call goes through the C ABI; there is
no Python or CPU implementation of the hot path behind it, and loading fails loudly if the library is missing."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpsat.so")

SAT, UNSAT, UNDEF = 0, 1, 2
DECIDE_REFERENCE, DECIDE_VSIDS = 0, 1
BCP_WATCHED, BCP_OCCURRENCE = 0, 1
STRATEGY_DISTRIBUTED, STRATEGY_UNIFORM, STRATEGY_SIMPLE = 0, 1, 2
E_NO_DEVICE = -2

RECORD_DTYPE = np.dtype([
    ("status", np.int32), ("reserved", np.int32), ("decisions", np.int64), ("implications", np.int64),
    ("conflicts", np.int64), ("learnt_clauses", np.int64), ("learnt_literals", np.int64), ("restarts", np.int64),
    ("watchers_visited", np.int64), ("clause_words_read", np.int64), ("learnt_hash", np.int64)])


class GpsatOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("decision", C.c_int32), ("bcp", C.c_int32),
                ("restart_first", C.c_int32), ("restart_factor", C.c_float), ("max_iterations", C.c_int32),
                ("stop_on_sat", C.c_int32), ("max_conflicts", C.c_int64), ("share_learnts", C.c_int32),
                ("share_max_len", C.c_int32), ("warps_per_block", C.c_int32), ("blocks", C.c_int32),
                ("arena_words", C.c_int64), ("dynamic_split", C.c_int32), ("split_gap", C.c_int32),
                ("split_burst", C.c_int32), ("share_import_max", C.c_int32), ("split_hand_words", C.c_int32),
                ("split_gap_hot", C.c_int32), ("split_at_start", C.c_int32), ("mesh_flags", C.c_int32), ("sweep_flags", C.c_int32),
                ("split_mode", C.c_int32), ("max_learnts", C.c_int32), ("split_min", C.c_int32), ("split_reserve", C.c_int32), ("phase_stats", C.c_int32), ("split_hard", C.c_int32)]


class GpsatStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("jobs_total", "jobs_done", "jobs_sat", "jobs_unsat", "jobs_undef", "decisions", "implications",
                 "conflicts", "learnt_clauses", "learnt_literals", "restarts", "watchers_visited",
                 "clause_words_read", "pool_clauses")] + \
               [("kernel_ms", C.c_double), ("kernel_launches", C.c_int32), ("blocks", C.c_int32),
                ("warps_per_block", C.c_int32), ("smem_bytes_per_block", C.c_int32), ("state_in_smem", C.c_int32),
                ("reserved", C.c_int32), ("splits", C.c_int64), ("warp_busy_frac", C.c_double),
                ("foreign_clauses", C.c_int64), ("steals", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class GpsatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gpsat error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `make lib` (python __graft_entry__.py build). "
                                "gpupsat_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.gpsat_last_error.restype = C.c_char_p
    L.gpsat_version.restype = C.c_char_p
    L.gpsat_cnf_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.gpsat_cnf_from_arrays.argtypes = [i64, vp, vp, C.POINTER(vp)]
    L.gpsat_cnf_free.argtypes = [vp]
    L.gpsat_cnf_preprocess.argtypes = [vp, C.POINTER(vp)]
    for name, rt in (("n_vars", i32), ("n_clauses", i64), ("n_lits", i64), ("offsets", vp), ("lits", vp),
                     ("status", i32), ("n_solved", i32), ("solved", vp), ("header_vars", i32),
                     ("header_clauses", i64), ("largest_clause", i32), ("most_common_var", i32),
                     ("most_common_freq", i32), ("n_lines", i32)):
        f = getattr(L, "gpsat_cnf_" + name)
        f.argtypes = [vp]
        f.restype = rt
    L.gpsat_choose_cubes.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), vp, i64]
    L.gpsat_opts_default.argtypes = [C.POINTER(GpsatOpts)]
    L.gpsat_create.argtypes = [C.POINTER(vp), i32, i64, vp, vp, C.POINTER(GpsatOpts)]
    L.gpsat_destroy.argtypes = [vp]
    L.gpsat_set_cubes.argtypes = [vp, i32, vp, vp]
    L.gpsat_propagate_all.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.gpsat_propagate.argtypes = [vp, i32, C.POINTER(i32), vp, C.POINTER(i32), C.POINTER(i64)]
    L.gpsat_eval_clauses.argtypes = [vp, i32, vp, vp, vp]
    L.gpsat_solve.argtypes = [vp, C.POINTER(i32), vp, C.POINTER(GpsatStats)]
    L.gpsat_job_records.argtypes = [vp, vp, i32]
    L.gpsat_last_kernel_ms.argtypes = [vp]
    L.gpsat_last_kernel_ms.restype = C.c_double
    L.gpsat_solve_begin.argtypes = [vp]
    L.gpsat_solve_step.argtypes = [vp, C.c_double, C.POINTER(i32), C.POINTER(i32)]
    L.gpsat_solve_end.argtypes = [vp, C.POINTER(i32), vp, C.POINTER(GpsatStats)]
    L.gpsat_request_stop.argtypes = [vp]
    L.gpsat_pool_export.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.gpsat_pool_import.argtypes = [vp, vp, i64]
    L.gpsat_exchange_block_words.argtypes = [i32]
    L.gpsat_exchange_block_words.restype = i64
    L.gpsat_exchange_pack.argtypes = [vp, vp, i64, i32, i32, i32]
    L.gpsat_exchange_unpack.argtypes = [vp, vp, i32, i32, i64, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                        C.POINTER(i64), C.POINTER(i64)]
    L.gpsat_debug_ctrl.argtypes = [vp, vp]
    L.gpsat_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.gpsat_mesh_export.argtypes = [vp, vp]
    L.gpsat_mesh_attach_ipc.argtypes = [vp, i32, i32, vp]
    L.gpsat_mesh_attach_local.argtypes = [C.POINTER(vp), i32]
    L.gpsat_mesh_detach.argtypes = [vp]
    L.gpsat_mesh_result_words.argtypes = [vp]
    L.gpsat_mesh_result_words.restype = i64
    L.gpsat_mesh_results_pack.argtypes = [vp, vp, i64]
    L.gpsat_mesh_results_unpack.argtypes = [vp, vp, i64, C.POINTER(i32), C.POINTER(GpsatStats)]
    L.gpsat_handle_device.argtypes = [vp]
    L.gpsat_debug_words.argtypes = [vp, vp, i32]
    L.gpsat_get_phase_stats.argtypes = [vp, vp]
    L.gpsat_multi_create.argtypes = [C.POINTER(vp), i32, vp, i32, i64, vp, vp, C.POINTER(GpsatOpts)]
    L.gpsat_multi_n_gpus.argtypes = [vp]
    L.gpsat_multi_set_time_limit.argtypes = [vp, C.c_double]
    L.gpsat_multi_set_cubes.argtypes = [vp, i32, vp, vp]
    L.gpsat_multi_solve.argtypes = [vp, C.POINTER(i32), vp, C.POINTER(GpsatStats), C.POINTER(i32)]
    L.gpsat_multi_job_records.argtypes = [vp, vp, i32]
    L.gpsat_multi_destroy.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise GpsatError(rc, lib().gpsat_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _copy(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()


class Cnf:
    """Host formula container (gpsat_cnf): ≙ FormulaData (FileManager/FormulaData.cuh:12-48)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    @classmethod
    def read(cls, path):
        out = C.c_void_p()
        _check(lib().gpsat_cnf_read(os.fsencode(path), C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_arrays(cls, offsets, lits):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lits = np.ascontiguousarray(lits, dtype=np.int32)
        out = C.c_void_p()
        _check(lib().gpsat_cnf_from_arrays(len(offsets) - 1, _p(offsets), _p(lits), C.byref(out)))
        return cls(out.value)

    def preprocess(self):
        out = C.c_void_p()
        _check(lib().gpsat_cnf_preprocess(self.h, C.byref(out)))
        return Cnf(out.value)

    def __del__(self):
        try:
            if self.h:
                lib().gpsat_cnf_free(self.h)
                self.h = None
        except Exception:
            pass

    n_vars = property(lambda s: lib().gpsat_cnf_n_vars(s.h))
    n_clauses = property(lambda s: lib().gpsat_cnf_n_clauses(s.h))
    n_lits = property(lambda s: lib().gpsat_cnf_n_lits(s.h))
    status = property(lambda s: lib().gpsat_cnf_status(s.h))
    header_vars = property(lambda s: lib().gpsat_cnf_header_vars(s.h))
    header_clauses = property(lambda s: lib().gpsat_cnf_header_clauses(s.h))
    largest_clause = property(lambda s: lib().gpsat_cnf_largest_clause(s.h))
    most_common_var = property(lambda s: lib().gpsat_cnf_most_common_var(s.h))
    most_common_freq = property(lambda s: lib().gpsat_cnf_most_common_freq(s.h))
    n_lines = property(lambda s: lib().gpsat_cnf_n_lines(s.h))

    @property
    def offsets(self):
        return _copy(lib().gpsat_cnf_offsets(self.h), self.n_clauses + 1, np.int64)

    @property
    def lits(self):
        return _copy(lib().gpsat_cnf_lits(self.h), self.n_lits, np.int32)

    @property
    def solved(self):
        return _copy(lib().gpsat_cnf_solved(self.h), lib().gpsat_cnf_n_solved(self.h), np.int32)

    def choose_cubes(self, blocks, threads, strategy=STRATEGY_DISTRIBUTED):
        """≙ MaxClauseJobChooser: returns an (n_cubes, k) int32 array of cube literals."""
        k, n = C.c_int32(0), C.c_int32(0)
        _check(lib().gpsat_choose_cubes(self.h, blocks, threads, strategy, C.byref(k), C.byref(n), None, 0))
        out = np.zeros(max(n.value * k.value, 1), dtype=np.int32)
        _check(lib().gpsat_choose_cubes(self.h, blocks, threads, strategy, C.byref(k), C.byref(n), _p(out),
                                        n.value * k.value))
        return out[: n.value * k.value].reshape(n.value, k.value)


def default_opts(**kw):
    o = GpsatOpts()
    lib().gpsat_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class Solver:
    """Device handle (gpsat_t): the formula resident in HBM plus the cube queue.
    ≙ DataToDevice + KernelContextStorage (SATSolver/DataToDevice.cuh:17-43, SATSolver/Parallelizer.cuh:18-29)."""

    def __init__(self, n_vars, offsets, lits, **opts):
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lits = np.ascontiguousarray(lits, dtype=np.int32)
        self.n_vars = int(n_vars)
        self.n_clauses = len(self.offsets) - 1
        self.opts = default_opts(**opts)
        h = C.c_void_p()
        _check(lib().gpsat_create(C.byref(h), self.n_vars, self.n_clauses, _p(self.offsets), _p(self.lits),
                                  C.byref(self.opts)))
        self.h = h
        self.n_cubes = 1

    def close(self):
        if getattr(self, "h", None):
            lib().gpsat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_cubes(self, cubes=None, cube_offsets=None, cube_lits=None):
        """cubes: (n, k) array, or explicit CSR (cube_offsets, cube_lits); None = the single empty cube."""
        self.empty_shard = False
        if cubes is not None:
            cubes = np.ascontiguousarray(cubes, dtype=np.int32)
            n, k = cubes.shape
            if n == 0:
                # an empty shard (more ranks than cubes) is NOT the single empty cube of sequential mode: this handle
                # has nothing to solve — solve_begin/step/end answer "done, UNSAT" without a launch
                self.empty_shard = True
                self.n_cubes = 0
                return
            cube_offsets = np.arange(0, n * k + 1, max(k, 1), dtype=np.int64) if k else np.zeros(n + 1, dtype=np.int64)
            cube_lits = cubes.reshape(-1)
        if cube_offsets is None:
            _check(lib().gpsat_set_cubes(self.h, 0, None, None))
            self.n_cubes = 1
            return
        co = np.ascontiguousarray(cube_offsets, dtype=np.int64)
        cl = np.ascontiguousarray(cube_lits, dtype=np.int32)
        _check(lib().gpsat_set_cubes(self.h, len(co) - 1, _p(co), _p(cl)))
        self.n_cubes = len(co) - 1

    def propagate_all(self, implied_stride=None, want_implied=True):
        n = self.n_cubes
        stride = self.n_vars if implied_stride is None else implied_stride
        status = np.zeros(n, dtype=np.int32)
        n_imp = np.zeros(n, dtype=np.int32)
        implied = np.full(max(n * stride, 1), -1, dtype=np.int32) if want_implied else None
        confl = np.zeros(n, dtype=np.int64)
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        _check(lib().gpsat_propagate_all(self.h, _p(status), _p(n_imp), _p(implied), stride, _p(confl), _p(rec)))
        return {"status": status, "n_implied": n_imp,
                "implied": implied[: n * stride].reshape(n, stride) if want_implied and stride else None,
                "conflict_clause": confl, "records": rec}

    def propagate(self, cube):
        st, n, cc = C.c_int32(0), C.c_int32(0), C.c_int64(-1)
        buf = np.zeros(max(self.n_vars, 1), dtype=np.int32)
        _check(lib().gpsat_propagate(self.h, cube, C.byref(st), _p(buf), C.byref(n), C.byref(cc)))
        return st.value, buf[: n.value].copy(), cc.value

    def eval_clauses(self, assignment, want_unit=True):
        a = np.ascontiguousarray(assignment, dtype=np.uint8).reshape(-1, self.n_vars)
        status = np.zeros((a.shape[0], self.n_clauses), dtype=np.int32)
        unit = np.zeros((a.shape[0], self.n_clauses), dtype=np.int32) if want_unit else None
        _check(lib().gpsat_eval_clauses(self.h, a.shape[0], _p(a), _p(status), _p(unit)))
        return status, unit

    def solve(self):
        verdict = C.c_int32(UNDEF)
        model = np.zeros(max(self.n_vars, 1), dtype=np.uint8)
        st = GpsatStats()
        _check(lib().gpsat_solve(self.h, C.byref(verdict), _p(model), C.byref(st)))
        return verdict.value, model[: self.n_vars], st.as_dict()

    def last_kernel_ms(self):
        return float(lib().gpsat_last_kernel_ms(self.h))

    def job_records(self, n=None):
        n = self.n_cubes if n is None else n
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        _check(lib().gpsat_job_records(self.h, _p(rec), n))
        return rec

    # --- mesh: several GPUs as one work pool over NVLink peer memory (include/gpsat.h: gpsat_mesh_*) ---
    def mesh_export(self):
        buf = np.zeros(64, dtype=np.uint8)
        _check(lib().gpsat_mesh_export(self.h, _p(buf)))
        return buf

    def mesh_attach_ipc(self, n_ranks, rank, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8).reshape(-1)
        assert handles.size == 64 * n_ranks
        _check(lib().gpsat_mesh_attach_ipc(self.h, n_ranks, rank, _p(handles)))

    def mesh_detach(self):
        _check(lib().gpsat_mesh_detach(self.h))

    def mesh_result_words(self):
        return int(lib().gpsat_mesh_result_words(self.h))

    def mesh_results_pack(self, block):
        _check(lib().gpsat_mesh_results_pack(self.h, C.c_void_p(block.data_ptr()), block.numel()))

    def mesh_results_unpack(self, block):
        verdict, st = C.c_int32(UNDEF), GpsatStats()
        _check(lib().gpsat_mesh_results_unpack(self.h, C.c_void_p(block.data_ptr()), block.numel(), C.byref(verdict),
                                               C.byref(st)))
        return verdict.value, st.as_dict()

    # --- epoch API (one process per GPU) ---
    def solve_begin(self):
        if getattr(self, "empty_shard", False):
            return
        _check(lib().gpsat_solve_begin(self.h))

    def solve_step(self, budget_ms=0.0):
        if getattr(self, "empty_shard", False):
            return True, UNSAT
        done, verdict = C.c_int32(0), C.c_int32(UNDEF)
        _check(lib().gpsat_solve_step(self.h, float(budget_ms), C.byref(done), C.byref(verdict)))
        return bool(done.value), verdict.value

    def solve_end(self):
        verdict = C.c_int32(UNDEF)
        model = np.zeros(max(self.n_vars, 1), dtype=np.uint8)
        st = GpsatStats()
        if getattr(self, "empty_shard", False):
            return UNSAT, model[: self.n_vars], st.as_dict()
        _check(lib().gpsat_solve_end(self.h, C.byref(verdict), _p(model), C.byref(st)))
        return verdict.value, model[: self.n_vars], st.as_dict()

    def debug_ctrl(self):
        out = np.zeros(16, dtype=np.int32)
        _check(lib().gpsat_debug_ctrl(self.h, _p(out)))
        return out

    def phase_stats(self):
        """gpsat_get_phase_stats: dict of arrays ns[8], count[8] + backtracked_levels, jobs, job_ns, idle_ns"""
        buf = np.zeros(20, dtype=np.int64)
        _check(lib().gpsat_get_phase_stats(self.h, _p(buf)))
        return {"ns": buf[:8].copy(), "count": buf[8:16].copy(), "backtracked_levels": int(buf[16]), "jobs": int(buf[17]),
                "job_ns": int(buf[18]), "idle_ns": int(buf[19])}

    def debug_words(self, n=160):
        out = np.zeros(n, dtype=np.int32)
        _check(lib().gpsat_debug_words(self.h, _p(out), n))
        return out

    def request_stop(self):
        _check(lib().gpsat_request_stop(self.h))

    def pool_export(self, cap_words=1 << 20):
        buf = np.zeros(cap_words, dtype=np.int32)
        n = C.c_int64(0)
        _check(lib().gpsat_pool_export(self.h, _p(buf), cap_words, C.byref(n)))
        return buf[: n.value].copy()

    def pool_import(self, words):
        words = np.ascontiguousarray(words, dtype=np.int32)
        _check(lib().gpsat_pool_import(self.h, _p(words), len(words)))

    # --- device-side epoch exchange: `block` / `gathered` are DEVICE buffers (anything with .data_ptr(), e.g. torch) ---
    def exchange_pack(self, block, rank, done, verdict):
        _check(lib().gpsat_exchange_pack(self.h, C.c_void_p(block.data_ptr()), block.numel(), rank, int(done), verdict))

    def exchange_unpack(self, gathered, n_ranks, rank):
        sat, done, undef = C.c_int32(-1), C.c_int32(0), C.c_int32(0)
        imported, jobs = C.c_int64(0), C.c_int64(0)
        _check(lib().gpsat_exchange_unpack(self.h, C.c_void_p(gathered.data_ptr()), n_ranks, rank,
                                           gathered.numel() // n_ranks, C.byref(sat), C.byref(done), C.byref(undef),
                                           C.byref(imported), C.byref(jobs)))
        return {"sat_rank": sat.value, "all_done": bool(done.value), "any_undef": bool(undef.value),
                "imported_clauses": imported.value, "jobs_done": jobs.value}


def mesh_attach_local(solvers):
    """In-process mesh over `solvers` (rank r = solvers[r]); every solver holds the same, complete cube list."""
    arr = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    _check(lib().gpsat_mesh_attach_local(arr, len(solvers)))


class MultiSolver:
    """gpsat_multi_*: N GPUs of this process as one solver (one host thread per GPU, mesh over NVLink peer memory)."""

    def __init__(self, n_vars, offsets, lits, n_gpus=0, devices=None, **opts):
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lits = np.ascontiguousarray(lits, dtype=np.int32)
        self.n_vars = int(n_vars)
        self.opts = default_opts(**opts)
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().gpsat_multi_create(C.byref(h), n_gpus, _p(dv), self.n_vars, len(self.offsets) - 1,
                                        _p(self.offsets), _p(self.lits), C.byref(self.opts)))
        self.h = h
        self.n_gpus = lib().gpsat_multi_n_gpus(h)
        self.n_cubes = 1

    def close(self):
        if getattr(self, "h", None):
            lib().gpsat_multi_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_time_limit(self, ms):
        _check(lib().gpsat_multi_set_time_limit(self.h, float(ms)))

    def set_cubes(self, cubes):
        if cubes is None:
            _check(lib().gpsat_multi_set_cubes(self.h, 0, None, None))
            self.n_cubes = 1
            return
        cubes = np.ascontiguousarray(cubes, dtype=np.int32)
        n, k = cubes.shape
        co = np.arange(0, n * k + 1, max(k, 1), dtype=np.int64) if k else np.zeros(n + 1, dtype=np.int64)
        cl = cubes.reshape(-1)
        _check(lib().gpsat_multi_set_cubes(self.h, n, _p(co), _p(cl)))
        self.n_cubes = n

    def solve(self):
        verdict, backend = C.c_int32(UNDEF), C.c_int32(0)
        model = np.zeros(max(self.n_vars, 1), dtype=np.uint8)
        st = GpsatStats()
        _check(lib().gpsat_multi_solve(self.h, C.byref(verdict), _p(model), C.byref(st), C.byref(backend)))
        d = st.as_dict()
        d["reduce_backend"] = "nccl" if backend.value == 1 else "host"
        return verdict.value, model[: self.n_vars], d

    def job_records(self):
        rec = np.zeros(self.n_cubes, dtype=RECORD_DTYPE)
        _check(lib().gpsat_multi_job_records(self.h, _p(rec), self.n_cubes))
        return rec


def solve_cnf(cnf: Cnf, blocks=32, threads=32, strategy=STRATEGY_DISTRIBUTED, sequential=False, **opts):
    """The reference's main() flow (SATSolver/main.cu:109-329) over the C ABI: preprocess, cubes, solve, merge model.
    Returns (verdict, model over the ORIGINAL variables or None, stats dict)."""
    pre = cnf.preprocess()
    n_vars = cnf.n_vars
    if pre.status != UNDEF:
        model = None
        if pre.status == SAT:
            model = np.ones(n_vars, dtype=np.uint8)
            for x in pre.solved:
                model[x >> 1] = x & 1
        return pre.status, model, {"solved_in_preprocessing": True}
    with Solver(n_vars, pre.offsets, pre.lits, **opts) as s:
        if sequential or n_vars < 3:
            s.set_cubes(None)
        else:
            s.set_cubes(pre.choose_cubes(blocks, threads, strategy))
        verdict, model, stats = s.solve()
    if verdict == SAT:
        model = model.copy()
        for x in pre.solved:
            model[x >> 1] = x & 1
    else:
        model = None
    return verdict, model, stats
