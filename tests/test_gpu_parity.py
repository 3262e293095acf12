"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same inputs — bit-exact."""
import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import check_model, pigeonhole, random_ksat
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu

FIELDS = [f for f in g.RECORD_DTYPE.names if f != "reserved"]


def _cmp_records(a, b, what):
    for f in FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert len(bad) == 0, f"{what}: field {f} differs at jobs {bad[:8]}: gpu {a[f][bad[:8]]} oracle {b[f][bad[:8]]}"


def _prep(offs, lits):
    cnf = g.Cnf.from_arrays(offs, lits)
    pre = cnf.preprocess()
    assert pre.status == g.UNDEF
    return cnf, pre


@pytest.mark.parametrize("seed", [0, 1])
def test_propagate_all_cubes_c2(seed):
    """config 2: n=250 m=1065, 4096 cubes: status, implied literal LISTS and conflict clause, all cubes."""
    offs, lits = random_ksat(250, 1065, seed)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(8, 32)
    assert cubes.shape == (4096, 12)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
        s.set_cubes(cubes)
        got = s.propagate_all()
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    co = np.arange(0, cubes.size + 1, 12, dtype=np.int64)
    want = o.run(co, cubes.reshape(-1), mode=1)
    assert np.array_equal(got["status"], want["records"]["status"])
    assert np.array_equal(got["n_implied"], want["n_implied"])
    assert np.array_equal(got["conflict_clause"], want["conflict_clause"])
    assert np.array_equal(got["implied"], want["implied"])
    _cmp_records(got["records"], want["records"], "propagate")


@pytest.mark.parametrize("n,m,seed,decision", [(20, 91, 0, 1), (20, 91, 1, 0), (50, 218, 0, 1), (50, 218, 1, 1),
                                               (50, 218, 2, 0), (100, 426, 0, 1), (150, 639, 0, 1)])
def test_sequential_solve_bit_exact(n, m, seed, decision):
    """config 1 style (-b 1 -t 1): one empty cube; every counter and the learnt-clause checksum equal the oracle's."""
    offs, lits = random_ksat(n, m, seed)
    cnf, pre = _prep(offs, lits)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, decision=decision, dynamic_split=0) as s:
        s.set_cubes(None)
        verdict, model, stats = s.solve()
        rec = s.job_records()
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    want = o.run(np.array([0, 0]), np.zeros(0), decision=decision)
    _cmp_records(rec, want["records"], "solve")
    assert verdict == want["records"]["status"][0]
    if verdict == g.SAT:
        assert np.array_equal(model, want["model"])
        assert check_model(pre.offsets, pre.lits, model)


def test_cube_solve_bit_exact_and_verdict():
    """cube-and-conquer on a uf100 instance with 2^7 cubes, no early stop: per-cube records equal the oracle's."""
    offs, lits = random_ksat(100, 426, 3)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(1, 8)          # 80 jobs wanted -> k = 7, 128 cubes
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0, dynamic_split=0) as s:
        s.set_cubes(cubes)
        verdict, model, stats = s.solve()
        rec = s.job_records()
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    k = cubes.shape[1]
    want = o.run(np.arange(0, cubes.size + 1, k, dtype=np.int64), cubes.reshape(-1), stop_on_sat=False)
    _cmp_records(rec, want["records"], "cube solve")
    any_sat = (want["records"]["status"] == 0).any()
    assert verdict == (g.SAT if any_sat else g.UNSAT)
    if verdict == g.SAT:
        assert check_model(pre.offsets, pre.lits, model)
    assert stats["jobs_done"] == len(cubes)


def test_pigeonhole_unsat():
    offs, lits = pigeonhole(7, 6)
    cnf, pre = _prep(offs, lits)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, dynamic_split=0) as s:
        s.set_cubes(None)
        verdict, _, _ = s.solve()
        rec = s.job_records()
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    want = o.run(np.array([0, 0]), np.zeros(0))
    assert verdict == g.UNSAT
    _cmp_records(rec, want["records"], "php(7,6)")


def test_eval_clauses():
    offs, lits = random_ksat(250, 1065, 5)
    rng = np.random.default_rng(0)
    assignment = rng.integers(0, 3, size=(7, 250)).astype(np.uint8)
    assignment[0, :] = 2
    with g.Solver(250, offs, lits) as s:
        st, unit = s.eval_clauses(assignment)
    o = Oracle(250, offs, lits)
    st2, unit2 = o.eval_clauses(assignment)
    assert np.array_equal(st, st2)
    assert np.array_equal(unit, unit2)


def test_dynamic_split_same_verdicts():
    """default options (dynamic splitting on): per-cube statuses equal the oracle's, every cube is closed, models verify"""
    for seed in (3, 6):
        offs, lits = random_ksat(120, 511, seed)
        cnf, pre = _prep(offs, lits)
        cubes = pre.choose_cubes(1, 2)          # few, long cubes: idle warps exist from the start -> splits happen
        with g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0) as s:
            s.set_cubes(cubes)
            verdict, model, stats = s.solve()
            rec = s.job_records()
        o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
        k = cubes.shape[1]
        want = o.run(np.arange(0, cubes.size + 1, k, dtype=np.int64), cubes.reshape(-1), stop_on_sat=False)
        assert np.array_equal(rec["status"], want["records"]["status"])
        assert stats["jobs_done"] == len(cubes)
        if verdict == g.SAT:
            assert check_model(pre.offsets, pre.lits, model)


def test_occurrence_bcp_small_with_conflicts():
    """GPSAT_BCP_OCCURRENCE (clause evaluation over occurrence lists, whole cube assigned first): status and implied
    SETS equal the oracle's batch propagation; a reported conflict clause is falsified under cube + implied."""
    from gpupsat_b200.instances import sweep_trails
    offs, lits = random_ksat(250, 1065, 7)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(8, 32)
    co = np.arange(0, cubes.size + 1, 12, dtype=np.int64)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, bcp=g.binding.BCP_OCCURRENCE) as s:
        s.set_cubes(cubes)
        got = s.propagate_all()
    want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(co, cubes.reshape(-1), mode=2)
    assert np.array_equal(got["status"], want["records"]["status"])
    for j in range(len(cubes)):
        mine = set(got["implied"][j, : got["n_implied"][j]].tolist())
        if got["status"][j] == g.UNDEF:
            assert mine == set(want["implied"][j, : want["n_implied"][j]].tolist())
        else:
            c = got["conflict_clause"][j]
            assert c >= 0
            trail = set(cubes[j].tolist()) | mine
            assert all((x ^ 1) in trail for x in pre.lits[pre.offsets[c]: pre.offsets[c + 1]])


def test_occurrence_bcp_mixed_clause_lengths():
    rng = np.random.default_rng(11)
    n = 400
    cl = []
    for _ in range(1500):
        ln = int(rng.choice([2, 3, 3, 4, 7]))
        vs = rng.choice(n, size=ln, replace=False)
        cl.append([int(2 * v + rng.integers(0, 2)) for v in vs])
    offs = np.cumsum([0] + [len(c) for c in cl]).astype(np.int64)
    lits = np.array([x for c in cl for x in c], dtype=np.int32)
    cubes = rng.integers(0, 2, size=(64, 10)).astype(np.int32) + 2 * rng.permuted(np.tile(np.arange(n), (64, 1)), axis=1)[:, :10].astype(np.int32)
    co = np.arange(0, cubes.size + 1, 10, dtype=np.int64)
    with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE) as s:
        s.set_cubes(cubes)
        got = s.propagate_all()
    want = Oracle(n, offs, lits).run(co, cubes.reshape(-1), mode=2)
    assert np.array_equal(got["status"], want["records"]["status"])
    for j in np.nonzero(got["status"] == g.UNDEF)[0]:
        assert set(got["implied"][j, : got["n_implied"][j]].tolist()) == set(want["implied"][j, : want["n_implied"][j]].tolist())


def _hub_3sat(n, m, hubs, hub_occ, seed):
    """pure 3-SAT with a few literals of very many occurrences (long occurrence lists: the bucket index's overflow
    path) on top of a uniform random background"""
    rng = np.random.default_rng(seed)
    cl = []
    for _ in range(m):
        vs = rng.choice(n, size=3, replace=False)
        cl.append([int(2 * v + rng.integers(0, 2)) for v in vs])
    for h in range(hubs):
        for _ in range(hub_occ):
            vs = rng.choice(np.arange(hubs, n), size=2, replace=False)
            c = [2 * h + (h & 1)] + [int(2 * v + rng.integers(0, 2)) for v in vs]
            cl.append([c[i] for i in rng.permutation(3)])
    offs = np.arange(0, 3 * len(cl) + 1, 3, dtype=np.int64)
    return offs, np.array([x for c in cl for x in c], dtype=np.int32)


@pytest.mark.parametrize("n,m,flags,hub_occ", [(600, 2200, 0, 40), (600, 2200, 1, 40), (3000, 12000, 1 | (10 << 8), 40),
                                               (3000, 12000, 0, 40), (600, 2200, 32, 40), (3000, 12000, 64, 40),
                                               (3000, 12000, 96, 40), (600, 2200, 64 | 2, 40), (3000, 12000, 0, 300)])
def test_occurrence_bcp_index_and_state_layouts(n, m, flags, hub_occ):
    """both large-database kernels — ternary state + bucket index (default for pure 3-SAT) and one CTA per job with the
    assigned-bit filter + global value fields (sweep_flags 1), the latter also with a filter SMALLER than the variable
    count (three variables per filter bit) — give the oracle's status and implied sets on an instance whose hub
    literals have 40 occurrences (bucket overflow path) and whose cubes falsify them; flags 32 / 64: the ternary
    kernel's lane-private code table / first hit kept during the scan"""
    offs, lits = _hub_3sat(n, m, 6, hub_occ, 5)     # 300: more than the 8-bit count field of a bucket holds
    rng = np.random.default_rng(6)
    J, K = 96, 24
    cubes = np.zeros((J, K), dtype=np.int32)
    for j in range(J):
        vs = rng.choice(np.arange(6, n), size=K - 6, replace=False)
        hub = [2 * h + (1 - (h & 1) if (j >> h) & 1 else int(rng.integers(0, 2))) for h in range(6)]   # often false hubs
        cubes[j] = np.array(hub + [int(2 * v + rng.integers(0, 2)) for v in vs], dtype=np.int32)
    cubes[3, 7] = cubes[3, 6]            # the same literal twice
    cubes[4, 9] = cubes[4, 8] ^ 1        # x and ~x: refuted without a clause to blame
    co = np.arange(0, cubes.size + 1, K, dtype=np.int64)
    with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE, sweep_flags=flags) as s:
        s.set_cubes(cubes)
        got = s.propagate_all()
    want = Oracle(n, offs, lits).run(co, cubes.reshape(-1), mode=2)
    assert got["status"][4] == g.UNSAT and got["conflict_clause"][4] == -1
    occ = np.bincount(lits, minlength=2 * n)
    assert occ.max() >= hub_occ
    assert np.array_equal(got["status"], want["records"]["status"])
    assert (got["status"] == g.UNSAT).any() and (got["status"] == g.UNDEF).any()
    for j in range(J):
        mine = set(got["implied"][j, : got["n_implied"][j]].tolist())
        if got["status"][j] == g.UNDEF:
            assert mine == set(want["implied"][j, : want["n_implied"][j]].tolist())
            # every occurrence of the negation of every trail literal was visited exactly once
            trail = np.concatenate([cubes[j], got["implied"][j, : got["n_implied"][j]]])
            assert got["records"]["watchers_visited"][j] == occ[trail ^ 1].sum()
        elif j != 4:
            c = got["conflict_clause"][j]
            assert c >= 0
            trail = set(cubes[j].tolist()) | mine
            assert all((x ^ 1) in trail for x in lits[offs[c]: offs[c + 1]])


def test_config4_large_database_sample():
    """config 4 at full size: n = 1e6, m = 4e6 planted 3-SAT, 32 jobs x 100k-literal trails; sample of jobs against
    the oracle (implied sets), all jobs conflict-free, implied literals agree with the planted assignment."""
    from gpupsat_b200.instances import planted_3sat_large, sweep_trails
    n, m, L, J = 1_000_000, 4_000_000, 100_000, 32
    offs, lits, planted = planted_3sat_large(n, m, 4)
    co, cl = sweep_trails(n, J, L, 4, planted)
    stride = 200_000
    with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE) as s:
        s.set_cubes(cube_offsets=co, cube_lits=cl)
        got = s.propagate_all(implied_stride=stride)
    assert (got["status"] == g.UNDEF).all()
    assert (got["n_implied"] > 1000).all() and (got["n_implied"] < stride).all()
    for j in range(J):
        imp = got["implied"][j, : got["n_implied"][j]]
        assert np.array_equal(planted[imp >> 1], imp & 1)
        assert len(np.unique(imp >> 1)) == len(imp)
    o = Oracle(n, offs, lits)
    for j in (0, 17):
        want = o.run(co[j: j + 2] - co[j], cl[co[j]: co[j + 1]], mode=2, implied_stride=stride)
        assert want["records"]["status"][0] == g.UNDEF
        assert set(got["implied"][j, : got["n_implied"][j]].tolist()) == set(want["implied"][0, : want["n_implied"][0]].tolist())


def test_budgeted_steps_and_device_exchange():
    """Two handles on one GPU play two ranks: cubes sharded j mod 2, budgeted steps (unfinished cubes park and resume),
    one exchange block per rank packed on the device, 'all-gathered' with torch.cat, unpacked into the other handle's
    foreign pool.  Every cube ends with the oracle's status; clauses crossed; every received clause is implied."""
    import torch
    from gpupsat_b200 import multi_gpu as mg
    offs, lits = random_ksat(150, 639, 2)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(1, 4)
    k = cubes.shape[1]
    want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(np.arange(0, cubes.size + 1, k, dtype=np.int64),
                                                          cubes.reshape(-1), stop_on_sat=False)
    words = mg.block_words(256)
    dev = torch.device("cuda", 0)
    solvers = [g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0, share_learnts=1, share_max_len=8)
               for _ in range(2)]
    blocks = [torch.zeros(words, dtype=torch.int32, device=dev) for _ in range(2)]
    for r, s in enumerate(solvers):
        s.set_cubes(mg.shard_cubes(cubes, r, 2))
        s.solve_begin()
    epochs, imported = 0, [0, 0]
    while True:
        for r, s in enumerate(solvers):
            done, verdict = s.solve_step(budget_ms=0.5)
            s.exchange_pack(blocks[r], r, done, verdict)
        gathered = torch.cat(blocks)
        torch.cuda.synchronize()
        infos = [s.exchange_unpack(gathered, 2, r) for r, s in enumerate(solvers)]
        assert infos[0]["all_done"] == infos[1]["all_done"] and infos[0]["sat_rank"] == infos[1]["sat_rank"]
        imported = [imported[r] + infos[r]["imported_clauses"] for r in range(2)]
        epochs += 1
        assert epochs < 5000
        if infos[0]["all_done"]:
            break
    g_host = gathered.cpu().numpy().reshape(2, -1)
    assert (g_host[:, 0] == mg.MAGIC).all()
    for r, s in enumerate(solvers):
        verdict, model, stats = s.solve_end()
        rec = s.job_records()
        assert np.array_equal(rec["status"], want["records"]["status"][r::2])
        assert stats["jobs_done"] == len(rec)
        assert stats["foreign_clauses"] == imported[r]
        if verdict == g.SAT:
            assert check_model(pre.offsets, pre.lits, model)
        s.close()
    assert epochs > 1 and sum(imported) > 0


def test_exchanged_clauses_are_implied():
    """every clause a handle exports (host-staged gpsat_pool_export) is refuted-when-negated by the oracle"""
    offs, lits = random_ksat(100, 426, 0)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(1, 2)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0, share_learnts=1, share_max_len=6) as s:
        s.set_cubes(cubes)
        s.solve()
        words = s.pool_export()
        assert len(s.pool_export()) == 0            # the export mark advanced
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    at, checked = 0, 0
    assert len(words) > 0
    while at < len(words) and checked < 40:
        ln = int(words[at])
        assert 1 <= ln <= 6
        neg = (words[at + 1: at + 1 + ln] ^ 1).astype(np.int32)
        assert o.run(np.array([0, ln], dtype=np.int64), neg)["records"]["status"][0] == g.UNSAT
        at += ln + 1
        checked += 1


def test_config5_pigeonhole_10_9_unsat():
    """config 5: PHP(10,9), 90 variables, 415 clauses, 4096 cubes — UNSAT (the reference as shipped answers UNDEF after
    1000 decisions; its cap-lifted host build needs ~7 minutes on one core for the same verdict)."""
    offs, lits = pigeonhole(10, 9)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(8, 32)
    assert cubes.shape == (4096, 12)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
        s.set_cubes(cubes)
        verdict, _, stats = s.solve()
        rec = s.job_records()
    assert verdict == g.UNSAT
    assert (rec["status"] == g.UNSAT).all() and stats["jobs_done"] == 4096
    # the same verdict from a different partition of the search space (k = 7) and from no partition at all
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
        s.set_cubes(pre.choose_cubes(1, 8))
        assert s.solve()[0] == g.UNSAT


def test_config2_and_larger_unsat_verdicts_and_cube_closure():
    """config 2 (uf250-1065 seed 0) and uf300-1278 seed 0: UNSAT, every one of the 4096 cubes closed; the per-cube
    statuses of a 1 % sample are cross-checked by the oracle solving those cubes on the CPU."""
    for n, m, seed, sample in ((250, 1065, 0, 40), (300, 1278, 0, 6)):
        offs, lits = random_ksat(n, m, seed)
        cnf, pre = _prep(offs, lits)
        cubes = pre.choose_cubes(8, 32)
        with g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0) as s:
            s.set_cubes(cubes)
            verdict, _, stats = s.solve()
            rec = s.job_records()
        assert verdict == g.UNSAT and stats["jobs_done"] == len(cubes)
        assert (rec["status"] == g.UNSAT).all()
        pick = np.linspace(0, len(cubes) - 1, sample).astype(int)
        k = cubes.shape[1]
        want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(np.arange(0, sample * k + 1, k, dtype=np.int64),
                                                              cubes[pick].reshape(-1), stop_on_sat=False)
        assert (want["records"]["status"] == g.UNSAT).all()


def test_sat_instance_model_through_cubes():
    """a satisfiable uf200 instance through 4096 cubes with early termination: model verifies against the CNF"""
    found = 0
    for seed in range(1, 40):
        offs, lits = random_ksat(200, 820, seed)          # r = 4.1: mostly satisfiable
        cnf, pre = _prep(offs, lits)
        with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
            s.set_cubes(pre.choose_cubes(8, 32))
            verdict, model, stats = s.solve()
        if verdict == g.SAT:
            assert check_model(pre.offsets, pre.lits, model)
            found += 1
            if found == 3:
                break
    assert found == 3


def test_edge_cases_ragged_cubes_long_clauses_absent_vars():
    """clauses longer than a warp, cubes of different lengths (empty, contradictory, touching variables that do not occur
    in the formula), both modes: every record equals the oracle's"""
    rng = np.random.default_rng(5)
    n = 96
    cl = []
    for _ in range(60):
        ln = int(rng.choice([2, 3, 3, 5, 40, 70]))
        vs = rng.choice(90, size=ln, replace=False)          # variables 90..95 never occur
        cl.append([int(2 * v + rng.integers(0, 2)) for v in vs])
    offs = np.cumsum([0] + [len(c) for c in cl]).astype(np.int64)
    lits = np.array([x for c in cl for x in c], dtype=np.int32)
    cubes = [[], [1], [1, 0], [3, 4, 9, 11, 20], [2 * v for v in range(30)], [2 * v + 1 for v in range(45)],
             [2 * 93 + 1, 5], [2 * 95, 2 * 95 + 1]]
    co = np.cumsum([0] + [len(c) for c in cubes]).astype(np.int64)
    clits = np.array([x for c in cubes for x in c], dtype=np.int32)
    o = Oracle(n, offs, lits)
    with g.Solver(n, offs, lits, stop_on_sat=0, dynamic_split=0) as s:
        s.set_cubes(cube_offsets=co, cube_lits=clits)
        got = s.propagate_all()
        want = o.run(co, clits, mode=1, stop_on_sat=False)
        _cmp_records(got["records"], want["records"], "ragged propagate")
        assert np.array_equal(got["implied"], want["implied"])
        verdict, model, stats = s.solve()
        rec = s.job_records()
    want = o.run(co, clits, mode=0, stop_on_sat=False)
    _cmp_records(rec, want["records"], "ragged solve")
    assert rec["status"][2] == g.UNSAT and rec["status"][7] == g.UNSAT      # x and ~x in one cube
    if verdict == g.SAT:
        assert check_model(offs, lits, model)


def test_maximum_cube_count_and_bad_arguments():
    """MAX_VARS = 15 -> 32768 cubes (Configs.cuh:40): all of them propagate; malformed inputs are refused, not run"""
    offs, lits = random_ksat(250, 1065, 1)
    cnf, pre = _prep(offs, lits)
    cubes = pre.choose_cubes(32, 1024)
    assert cubes.shape == (32768, 15)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
        s.set_cubes(cubes)
        got = s.propagate_all(want_implied=False)
        assert set(np.unique(got["status"]).tolist()) <= {g.UNSAT, g.UNDEF}
        want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(np.arange(0, 15 * 512 + 1, 15, dtype=np.int64),
                                                              cubes[:512].reshape(-1), mode=1)
        assert np.array_equal(got["status"][:512], want["records"]["status"])
        assert np.array_equal(got["n_implied"][:512], want["n_implied"])
        with pytest.raises(g.GpsatError):
            s.set_cubes(np.array([[2 * cnf.n_vars + 4]], dtype=np.int32))          # literal out of range
    with pytest.raises(g.GpsatError):
        g.Solver(3, np.array([0, 1, 3], dtype=np.int64), np.array([1, 2, 4], dtype=np.int32))   # unit clause


def test_state_in_global_memory_path():
    """formulas whose per-job state does not fit in shared memory (n = 4000) run the <global state> kernel variant:
    sequential solve bit-exact against the oracle, cube solve returns a verified model"""
    n, m = 4000, 12000                                  # r = 3.0: satisfiable, solved mostly by propagation
    offs, lits = random_ksat(n, m, 2)
    cnf, pre = _prep(offs, lits)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, dynamic_split=0) as s:
        s.set_cubes(None)
        verdict, model, stats = s.solve()
        rec = s.job_records()
        assert stats["state_in_smem"] == 0
    want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(np.array([0, 0]), np.zeros(0))
    _cmp_records(rec, want["records"], "global-state solve")
    assert verdict == g.SAT and check_model(pre.offsets, pre.lits, model)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits) as s:
        s.set_cubes(pre.choose_cubes(2, 32))
        verdict, model, stats = s.solve()
    assert verdict == g.SAT and check_model(pre.offsets, pre.lits, model)


# ---- pins against outputs of the REFERENCE itself (tests/golden/reference_outputs.json, generated by executing
# oracle/_ref = the reference's own classes; tests/golden/make_golden.py) run through the CUDA path ----------------------
from tests.helpers import golden  # noqa: E402

G = golden()


@pytest.mark.parametrize("name", sorted(k for k in G["verdicts"] if k.startswith(("uf20", "uf50", "php"))))
def test_golden_reference_verdicts_through_the_cuda_path(name):
    """the 20 uf20/uf50 verdicts and the 3 pigeonhole verdicts the reference produced, sequential mode (-b 1 -t 1)"""
    if name.startswith("uf"):
        n, m = int(name.split("-")[0][2:]), int(name.split("-")[1])
        offs, lits = random_ksat(n, m, int(name.split("seed")[1]))
    else:
        _, p, h = name.split("-")
        offs, lits = pigeonhole(int(p), int(h))
    cnf = g.Cnf.from_arrays(offs, lits)
    verdict, model, stats = g.solve_cnf(cnf, sequential=True)
    assert verdict == G["verdicts"][name]
    if verdict == g.SAT:
        assert check_model(offs, lits, model)


def test_max_iterations_1000_reproduces_the_reference_undef():
    """as shipped the reference gives up after 1000 decisions (SATSolver/Configs.cuh:23): uf100-426 seed 0 -> UNDEF;
    the same cap on the CUDA path (reference decision rule) gives UNDEF too, no cap gives the cap-lifted UNSAT"""
    offs, lits = random_ksat(100, 426, 0)
    cnf = g.Cnf.from_arrays(offs, lits)
    assert G["verdicts"]["uf100-426-seed0-as-shipped"] == g.UNDEF and G["verdicts"]["uf100-426-seed0-nocap"] == g.UNSAT
    v_cap, _, _ = g.solve_cnf(cnf, sequential=True, decision=g.DECIDE_REFERENCE, max_iterations=1000, dynamic_split=0)
    v_free, _, _ = g.solve_cnf(cnf, sequential=True, decision=g.DECIDE_REFERENCE, dynamic_split=0)
    assert v_cap == g.UNDEF and v_free == g.UNSAT


@pytest.mark.parametrize("name,n,m,seed", [("uf50-218-seed0", 50, 218, 0), ("uf250-1065-seed0", 250, 1065, 0)])
def test_eval_clauses_equals_reference_clause_status(name, n, m, seed):
    """gpsat_eval_clauses against VariablesStateHandler::clause_status executed by the reference (golden fixture)"""
    offs, lits = random_ksat(n, m, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    rows = G["clause_status"][name]
    a = np.array([r["assignment"] for r in rows], dtype=np.uint8)
    with g.Solver(n, pre.offsets, pre.lits) as s:
        st, unit = s.eval_clauses(a)
    assert st.tolist() == [r["status"] for r in rows]
    assert unit.tolist() == [r["unit"] for r in rows]


def test_single_cube_propagate_equals_propagate_all_in_ternary_mode():
    """gpsat_propagate(j) narrows the job list to one cube; the ternary sweep kernel's per-cube info (short-list count,
    distinct-variables flag) must follow it (it read cube 0's)"""
    offs, lits = random_ksat(600, 2400, 11)
    rng = np.random.default_rng(3)
    cubes = []
    for j in range(12):
        ln = int(rng.integers(5, 120))
        vs = rng.choice(600, size=ln, replace=(j % 3 == 0))          # every third cube repeats variables
        cubes.append((2 * vs + rng.integers(0, 2, size=ln)).astype(np.int32))
    co = np.concatenate([[0], np.cumsum([len(c) for c in cubes])]).astype(np.int64)
    cl = np.concatenate(cubes)
    with g.Solver(600, offs, lits, bcp=g.binding.BCP_OCCURRENCE) as s:
        s.set_cubes(cube_offsets=co, cube_lits=cl)
        allr = s.propagate_all()
        for j in range(len(cubes)):
            st, imp, cc = s.propagate(j)
            assert st == allr["status"][j]
            if st == g.UNDEF:
                assert set(imp.tolist()) == set(allr["implied"][j, : allr["n_implied"][j]].tolist())


def test_cubes_longer_than_the_split_buffer_are_solved_in_place():
    """a cube of more than GPSAT_DQ_MAXK - 1 = 63 literals cannot be copied into the warp's cube area (it
    used to overrun it): it runs in place, without splitting or parking, and gets the oracle's status"""
    offs, lits = random_ksat(300, 1150, 8)                  # under-constrained: long consistent cubes exist
    cnf, pre = _prep(offs, lits)
    rng = np.random.default_rng(5)
    cubes, co = [], [0]
    for j in range(24):
        ln = int(rng.integers(60, 140))
        vs = rng.choice(300, size=ln, replace=False)
        c = (2 * vs + rng.integers(0, 2, size=ln)).astype(np.int32)
        cubes.append(c)
        co.append(co[-1] + ln)
    co = np.array(co, dtype=np.int64)
    cl = np.concatenate(cubes)
    want = Oracle(cnf.n_vars, pre.offsets, pre.lits).run(co, cl, stop_on_sat=False)
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, stop_on_sat=0) as s:      # dynamic_split on (the default)
        s.set_cubes(cube_offsets=co, cube_lits=cl)
        verdict, model, stats = s.solve()
        rec = s.job_records()
    assert np.array_equal(rec["status"], want["records"]["status"])
    assert stats["jobs_done"] == 24
