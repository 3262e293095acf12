"""TEST-ONLY: builds and binds the lockstep emulator of the kernel source (tests/emu/emu.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libgpsat_emu.so")

RECORD_DTYPE = np.dtype([
    ("status", np.int32), ("reserved", np.int32), ("decisions", np.int64), ("implications", np.int64),
    ("conflicts", np.int64), ("learnt_clauses", np.int64), ("learnt_literals", np.int64), ("restarts", np.int64),
    ("watchers_visited", np.int64), ("clause_words_read", np.int64), ("learnt_hash", np.int64)])


class SolveParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("decision", C.c_int32), ("bcp", C.c_int32), ("restart_first", C.c_int32),
                ("restart_factor", C.c_float), ("max_iterations", C.c_int32), ("stop_on_sat", C.c_int32),
                ("share_learnts", C.c_int32), ("share_max_len", C.c_int32), ("max_learnts_first", C.c_int32),
                ("learnt_refs_cap", C.c_int32), ("max_conflicts", C.c_int64), ("arena_words", C.c_int64),
                ("implied_stride", C.c_int64), ("dynamic_split", C.c_int32), ("split_force", C.c_int32),
                ("split_gap", C.c_int32), ("split_burst", C.c_int32), ("share_import_max", C.c_int32),
                ("split_gap_hot", C.c_int32), ("split_hot_demand", C.c_int32), ("split_at_start", C.c_int32),
                ("mesh_flags", C.c_int32), ("split_reserve", C.c_int32), ("phase_stats", C.c_int32), ("split_mode", C.c_int32), ("split_min", C.c_int32), ("split_hard", C.c_int32)]


def build(force=False):
    srcs = [os.path.join(HERE, "emu.cpp"), os.path.join(ROOT, "gpupsat_b200", "csrc", "host_formula.cpp")]
    deps = srcs + [os.path.join(ROOT, "gpupsat_b200", "csrc", f) for f in
                   ("cdcl_warp.inl", "warp_lockstep.h", "gpsat_device.h", "host_formula.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-fPIC", "-shared", "-o", SO] + srcs)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(n_vars, offsets, lits, cube_offsets, cube_lits, *, mode=0, decision=1, restart_first=100, restart_factor=1.3,
        max_iterations=0, max_conflicts=0, max_learnts_first=None, learnt_refs_cap=16384, arena_words=1 << 19,
        stop_on_sat=True, share_learnts=0, share_max_len=8, pool=None, pool_cursor=None, dynamic_split=0,
        split_force=0, split_gap=8, split_burst=4, budget_ticks=0, share_import_max=256, packed=False):
    """packed: the kernel variant that runs beside a formula staged in shared memory (one word per cl2 / occ2 pair,
    16-bit level / trail / trail_lim)"""
    build()
    lib = C.CDLL(SO)
    lib.gpsat_emu_set_packed(C.c_int(1 if packed else 0))
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lits = np.ascontiguousarray(lits, dtype=np.int32)
    co = np.ascontiguousarray(cube_offsets, dtype=np.int64)
    cl = np.ascontiguousarray(cube_lits, dtype=np.int32)
    n_cubes = len(co) - 1
    m = len(offsets) - 1
    if max_learnts_first is None:
        max_learnts_first = max(min(max(m // 3, 300), learnt_refs_cap - n_vars - 2), 1)
    P = SolveParams(mode, decision, 0, restart_first, restart_factor, max_iterations, 1 if stop_on_sat else 0,
                    share_learnts, share_max_len, max_learnts_first, learnt_refs_cap, max_conflicts, arena_words,
                    n_vars, dynamic_split, split_force, split_gap, split_burst, share_import_max)
    rec = np.zeros(n_cubes, dtype=RECORD_DTYPE)
    model = np.zeros(max(n_vars, 1), dtype=np.uint8)
    sat_job = C.c_int32(-1)
    launches = C.c_int32(0)
    implied = np.full(max(n_cubes * n_vars, 1), -1, dtype=np.int32)
    n_implied = np.zeros(n_cubes, dtype=np.int32)
    confl = np.full(n_cubes, -1, dtype=np.int64)
    if pool is None:
        pool = np.zeros(1 << 16, dtype=np.int32)
        pool_cursor = np.zeros(4, dtype=np.int32)
    rc = lib.gpsat_emu_run(C.c_int32(n_vars), C.c_int64(m), _p(offsets), _p(lits), C.byref(P), C.c_int32(n_cubes),
                           _p(co), _p(cl), _p(rec), _p(model), C.byref(sat_job), _p(implied), _p(n_implied), _p(confl),
                           _p(pool), _p(pool_cursor), C.c_int32(len(pool)), C.c_uint64(budget_ticks),
                           C.byref(launches))
    assert rc == 0, rc
    return {"launches": launches.value, "records": rec, "sat_job": sat_job.value, "model": model[:n_vars],
            "implied": implied.reshape(n_cubes, n_vars) if n_vars else implied, "n_implied": n_implied,
            "conflict_clause": confl, "pool": pool, "pool_cursor": pool_cursor}


def bucket_index(n_vars, offsets, lits):
    """host-side bucket index of the ternary sweep kernel (host_formula.cpp: build_sweep_index):
    (bucket words [2n+2, 16], orange [2n, 2]) or None when the database does not qualify"""
    build()
    lib = C.CDLL(SO)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lits = np.ascontiguousarray(lits, dtype=np.int32)
    bucket = np.zeros((2 * n_vars + 2, 16), dtype=np.uint32)
    orange = np.zeros((2 * n_vars, 2), dtype=np.int32)
    n_entries = C.c_int64(0)
    rc = lib.gpsat_emu_bucket_index(C.c_int32(n_vars), C.c_int64(len(offsets) - 1), _p(offsets), _p(lits), _p(bucket),
                                    _p(orange), C.byref(n_entries))
    if rc == 1:
        return None
    assert rc == 0
    return bucket, orange


def order_cubes(n_vars, offsets, lits, cube_offsets, cube_lits):
    """host_formula.cpp: order_cubes_for_sweep -> (sorted literals, info word per cube)"""
    build()
    lib = C.CDLL(SO)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lits = np.ascontiguousarray(lits, dtype=np.int32)
    co = np.ascontiguousarray(cube_offsets, dtype=np.int64)
    cl = np.ascontiguousarray(cube_lits, dtype=np.int32)
    out = np.zeros(max(int(co[-1] - co[0]), 1), dtype=np.int32)
    info = np.zeros(len(co) - 1, dtype=np.int32)
    rc = lib.gpsat_emu_order_cubes(C.c_int32(n_vars), C.c_int64(len(offsets) - 1), _p(offsets), _p(lits),
                                   C.c_int32(len(co) - 1), _p(co), _p(cl), _p(out), _p(info))
    assert rc == 0
    return out[: int(co[-1] - co[0])], info


def plan(n_vars, n_lits, n_clauses, *, phase_stats=0, solve=1, warps=0, w_max=28, w_auto_max=24, smem_max=232448):
    """gpsat_plan_warps: launch geometry of the CDCL kernel for a formula of this size on an SM with smem_max bytes"""
    build()
    lib = C.CDLL(SO)
    out = np.zeros(8, dtype=np.int64)
    lib.gpsat_emu_plan(C.c_int32(n_vars), C.c_int64(n_lits), C.c_int64(n_clauses), C.c_int32(phase_stats), C.c_int32(solve),
                       C.c_int32(warps), C.c_int32(w_max), C.c_int32(w_auto_max), C.c_int64(smem_max), _p(out))
    keys = ("warps", "state_in_smem", "formula_in_smem", "formula_smem_words", "smem_bytes", "state_words", "idx16", "lbuf_words")
    return dict(zip(keys, (int(x) for x in out)))
