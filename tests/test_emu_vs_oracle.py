"""The kernel source (gpupsat_b200/csrc/cdcl_warp.inl) compiled in the test-only lockstep mode and stepped on the
CPU, against the oracle: every per-job counter, the learnt-clause checksum, models, implied lists.  This is the
CPU-side guard for the warp program itself; the -m gpu tests run the same comparisons on the real kernel."""
import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import check_model, pigeonhole, random_ksat
from oracle.binding import Oracle
from tests.emu import binding as emu
from tests.helpers import cube_csr

EMPTY = (np.array([0, 0], dtype=np.int64), np.zeros(0, dtype=np.int32))
FIELDS = [f for f in emu.RECORD_DTYPE.names if f != "reserved"]


def same(a, b):
    return [f for f in FIELDS if not np.array_equal(a[f], b[f])]


@pytest.mark.parametrize("n,m,seed", [(20, 91, 0), (20, 91, 1), (50, 218, 0), (50, 218, 3), (100, 426, 0), (120, 511, 2)])
@pytest.mark.parametrize("decision", [0, 1])
def test_sequential_solve(n, m, seed, decision):
    offs, lits = random_ksat(n, m, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    o = Oracle(n, pre.offsets, pre.lits)
    a = o.run(*EMPTY, decision=decision, max_conflicts=60000)
    b = emu.run(n, pre.offsets, pre.lits, *EMPTY, decision=decision, max_conflicts=60000)
    assert same(a["records"], b["records"]) == []
    if a["records"]["status"][0] == g.SAT:
        assert np.array_equal(a["model"], b["model"]) and check_model(pre.offsets, pre.lits, b["model"])


def test_reduce_db_and_restarts_exercised():
    offs, lits = random_ksat(150, 639, 0)
    o = Oracle(150, offs, lits)
    a = o.run(*EMPTY, max_learnts_first=200)
    b = emu.run(150, offs, lits, *EMPTY, max_learnts_first=200)
    assert a["records"]["restarts"][0] > 3 and a["records"]["learnt_clauses"][0] > 1000
    assert same(a["records"], b["records"]) == []


def test_tiny_arena_reports_oom_identically():
    offs, lits = random_ksat(100, 426, 0)
    kw = dict(arena_words=6 * 100 + 64 + 512 + 3000, learnt_refs_cap=512, max_learnts_first=400)
    a = Oracle(100, offs, lits).run(*EMPTY, **kw)
    b = emu.run(100, offs, lits, *EMPTY, **kw)
    assert same(a["records"], b["records"]) == []


@pytest.mark.parametrize("p,h", [(5, 4), (7, 6)])
def test_pigeonhole(p, h):
    offs, lits = pigeonhole(p, h)
    n = p * h
    a = Oracle(n, offs, lits).run(*EMPTY)
    b = emu.run(n, offs, lits, *EMPTY)
    assert a["records"]["status"][0] == g.UNSAT
    assert same(a["records"], b["records"]) == []


def test_cube_solve_and_propagate():
    offs, lits = random_ksat(100, 426, 3)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cubes = pre.choose_cubes(1, 8)
    co, cl = cube_csr(cubes)
    o = Oracle(100, pre.offsets, pre.lits)
    a = o.run(co, cl, stop_on_sat=False)
    b = emu.run(100, pre.offsets, pre.lits, co, cl, stop_on_sat=False)
    assert same(a["records"], b["records"]) == []
    a = o.run(co, cl, mode=1)
    b = emu.run(100, pre.offsets, pre.lits, co, cl, mode=1)
    assert same(a["records"], b["records"]) == []
    assert np.array_equal(a["implied"], b["implied"]) and np.array_equal(a["n_implied"], b["n_implied"])
    assert np.array_equal(a["conflict_clause"], b["conflict_clause"])


def test_long_clauses_and_ragged_cubes():
    # clauses longer than a warp (chunked scans), cubes of different lengths, an empty cube, a contradictory cube
    rng = np.random.default_rng(5)
    n = 90
    cl = []
    for _ in range(60):
        ln = int(rng.choice([2, 3, 3, 5, 40, 70]))
        vs = rng.choice(n, size=ln, replace=False)
        cl.append([int(2 * v + rng.integers(0, 2)) for v in vs])
    offs = np.cumsum([0] + [len(c) for c in cl]).astype(np.int64)
    lits = np.array([x for c in cl for x in c], dtype=np.int32)
    cubes = [[], [1], [1, 0], [3, 4, 9, 11, 20], [2 * v for v in range(30)], [2 * v + 1 for v in range(45)]]
    co = np.cumsum([0] + [len(c) for c in cubes]).astype(np.int64)
    clits = np.array([x for c in cubes for x in c], dtype=np.int32)
    o = Oracle(n, offs, lits)
    for mode in (1, 0):
        a = o.run(co, clits, mode=mode, stop_on_sat=False)
        b = emu.run(n, offs, lits, co, clits, mode=mode, stop_on_sat=False)
        assert same(a["records"], b["records"]) == []
    assert a["records"]["status"][2] == g.UNSAT          # x0 and ~x0 in one cube


def test_pool_import_and_publish():
    """shared learnt pool: clauses published by one run are imported by the next; verdicts unchanged and every
    pool record is a clause implied by the formula (checked by refuting formula + negated clause)."""
    offs, lits = random_ksat(60, 258, 1)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cubes = pre.choose_cubes(1, 2)
    co, cl = cube_csr(cubes)
    base = emu.run(60, pre.offsets, pre.lits, co, cl, stop_on_sat=False)
    first = emu.run(60, pre.offsets, pre.lits, co, cl, stop_on_sat=False, share_learnts=1, share_max_len=6)
    assert np.array_equal(base["records"]["status"], first["records"]["status"])
    slots = first["pool"][: 16 * first["pool_cursor"][0]].reshape(-1, 16)   # fixed 16-word slots [len, lit0 ...]
    assert first["pool_cursor"][1] > 0 and first["pool_cursor"][1] == first["pool_cursor"][0]
    o = Oracle(60, pre.offsets, pre.lits)
    for rec in slots[:25]:
        ln = int(rec[0])
        assert 1 <= ln <= 6
        neg = np.array([x ^ 1 for x in rec[1: 1 + ln]], dtype=np.int32)   # formula AND not(clause) must be UNSAT
        r = o.run(np.array([0, ln], dtype=np.int64), neg)
        assert r["records"]["status"][0] == g.UNSAT
    again = emu.run(60, pre.offsets, pre.lits, co, cl, stop_on_sat=False, share_learnts=1, share_max_len=6,
                    pool=first["pool"], pool_cursor=first["pool_cursor"])
    assert np.array_equal(base["records"]["status"], again["records"]["status"])
    assert again["records"]["conflicts"].sum() <= base["records"]["conflicts"].sum()


def test_dynamic_split_keeps_verdicts():
    """forced splitting at every restart (test hook): children are queued and solved; every original cube gets the
    same status as without splitting, SAT models verify, and all queued children are closed."""
    for n, m, seed, bt in ((100, 426, 0, (1, 1)), (100, 426, 3, (1, 8)), (120, 511, 1, (1, 2))):
        offs, lits = random_ksat(n, m, seed)
        pre = g.Cnf.from_arrays(offs, lits).preprocess()
        cubes = pre.choose_cubes(*bt)
        co, cl = cube_csr(cubes)
        base = emu.run(n, pre.offsets, pre.lits, co, cl, stop_on_sat=False)
        split = emu.run(n, pre.offsets, pre.lits, co, cl, stop_on_sat=False, dynamic_split=1, split_force=1)
        assert np.array_equal(base["records"]["status"], split["records"]["status"])
        if bt == (1, 1):
            assert split["records"]["reserved"].sum() > 0      # cubes long enough to reach a restart were split
        if split["sat_job"] >= 0:
            assert check_model(pre.offsets, pre.lits, split["model"])
        # without the hook (nobody is ever idle in the one-warp emulator) the run is identical to no splitting
        quiet_split = emu.run(n, pre.offsets, pre.lits, co, cl, stop_on_sat=False, dynamic_split=1)
        assert same(base["records"], quiet_split["records"]) == []


def test_budgeted_steps_suspend_and_resume():
    """gpsat_solve_step with a budget: a cube that outlives the step parks itself in the ring (cube, VSIDS counters,
    level-0 facts, newest learnt clauses) and a later launch resumes it; verdicts are those of the unbudgeted run."""
    for n, m, seed, bt in ((100, 426, 0, (1, 1)), (120, 511, 1, (1, 2)), (100, 426, 3, (1, 8))):
        offs, lits = random_ksat(n, m, seed)
        pre = g.Cnf.from_arrays(offs, lits).preprocess()
        cubes = pre.choose_cubes(*bt)
        co, cl = cube_csr(cubes)
        base = emu.run(n, pre.offsets, pre.lits, co, cl, stop_on_sat=False)
        stepped = emu.run(n, pre.offsets, pre.lits, co, cl, stop_on_sat=False, dynamic_split=1, budget_ticks=3)
        assert np.array_equal(base["records"]["status"], stepped["records"]["status"])
        if base["records"]["conflicts"].max() > 64:
            assert stepped["launches"] > 1          # at least one job was parked and resumed by a later launch
        if stepped["sat_job"] >= 0:
            assert check_model(pre.offsets, pre.lits, stepped["model"])


# ---- the variant that runs beside a formula staged in shared memory ---------------------------------------------------
# (WarpSolverT<true>: cl2 / occ2 pairs packed into one word, 16-bit level / trail / trail_lim; kernels.cu stages the
# packed copy, plan_geometry picks the 16-bit layout).  Same program, other element types: every counter must still be
# the oracle's.
@pytest.mark.parametrize("n,m,seed", [(20, 91, 1), (50, 218, 3), (100, 426, 0), (120, 511, 2), (150, 639, 0)])
def test_packed_variant_sequential_solve(n, m, seed):
    offs, lits = random_ksat(n, m, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    kw = dict(max_learnts_first=200) if n == 150 else dict(max_conflicts=60000)
    a = Oracle(n, pre.offsets, pre.lits).run(*EMPTY, **kw)
    b = emu.run(n, pre.offsets, pre.lits, *EMPTY, packed=True, **kw)
    assert same(a["records"], b["records"]) == []
    if a["records"]["status"][0] == g.SAT:
        assert np.array_equal(a["model"], b["model"]) and check_model(pre.offsets, pre.lits, b["model"])


def test_packed_variant_cubes_long_clauses_split_and_steps():
    # cubes: solve and propagate, implied lists and conflict clauses
    offs, lits = random_ksat(100, 426, 3)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    co, cl = cube_csr(pre.choose_cubes(1, 8))
    o = Oracle(100, pre.offsets, pre.lits)
    for mode in (0, 1):
        a = o.run(co, cl, mode=mode, stop_on_sat=False)
        b = emu.run(100, pre.offsets, pre.lits, co, cl, mode=mode, stop_on_sat=False, packed=True)
        assert same(a["records"], b["records"]) == []
    assert np.array_equal(a["implied"], b["implied"]) and np.array_equal(a["conflict_clause"], b["conflict_clause"])
    # clauses longer than a warp, ragged cubes
    rng = np.random.default_rng(5)
    n = 90
    cls = []
    for _ in range(60):
        ln = int(rng.choice([2, 3, 3, 5, 40, 70]))
        cls.append([int(2 * v + rng.integers(0, 2)) for v in rng.choice(n, size=ln, replace=False)])
    offs2 = np.cumsum([0] + [len(c) for c in cls]).astype(np.int64)
    lits2 = np.array([x for c in cls for x in c], dtype=np.int32)
    cubes = [[], [1], [1, 0], [3, 4, 9, 11, 20], [2 * v for v in range(30)], [2 * v + 1 for v in range(45)]]
    co2 = np.cumsum([0] + [len(c) for c in cubes]).astype(np.int64)
    cl2 = np.array([x for c in cubes for x in c], dtype=np.int32)
    for mode in (1, 0):
        a = Oracle(n, offs2, lits2).run(co2, cl2, mode=mode, stop_on_sat=False)
        b = emu.run(n, offs2, lits2, co2, cl2, mode=mode, stop_on_sat=False, packed=True)
        assert same(a["records"], b["records"]) == []
    # forced splitting and budgeted steps: same statuses as the plain run of the unpacked variant
    base = emu.run(100, pre.offsets, pre.lits, co, cl, stop_on_sat=False)
    split = emu.run(100, pre.offsets, pre.lits, co, cl, stop_on_sat=False, dynamic_split=1, split_force=1, packed=True)
    stepped = emu.run(100, pre.offsets, pre.lits, co, cl, stop_on_sat=False, dynamic_split=1, budget_ticks=3, packed=True)
    assert np.array_equal(base["records"]["status"], split["records"]["status"])
    assert np.array_equal(base["records"]["status"], stepped["records"]["status"])
