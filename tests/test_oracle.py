"""The CPU oracle (oracle/gpsat_oracle.cpp) pinned against the reference: golden fixtures made by running the
reference's own code (always), and the reference host build itself where present (this container)."""
import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import check_model, parse_dimacs_text, pigeonhole, random_ksat
from oracle.binding import Oracle, Reference
from tests.helpers import cube_csr, golden, have_ref, model_from_lits

G = golden()
EMPTY = (np.array([0, 0], dtype=np.int64), np.zeros(0, dtype=np.int32))


def oracle_verdict(offs, lits, **kw):
    cnf = g.Cnf.from_arrays(offs, lits)
    pre = cnf.preprocess()
    if pre.status != g.UNDEF:
        return pre.status, None, cnf, pre
    o = Oracle(cnf.n_vars, pre.offsets, pre.lits)
    r = o.run(*EMPTY, **kw)
    model = None
    if r["records"]["status"][0] == g.SAT:
        model = r["model"].copy()
        for x in pre.solved:
            model[x >> 1] = x & 1
    return int(r["records"]["status"][0]), model, cnf, pre


@pytest.mark.parametrize("name", sorted(G["tests_cnf"]))
@pytest.mark.parametrize("decision", [0, 1])
def test_reference_cnf_fixtures(name, decision):
    d = G["tests_cnf"][name]
    offs, lits, _ = parse_dimacs_text(d["dimacs"])
    verdict, model, cnf, pre = oracle_verdict(offs, lits, decision=decision)
    assert {0: "SAT", 1: "UNSAT"}[verdict] == d["expected"]
    assert verdict == d["verdict"]
    if verdict == g.SAT:
        assert check_model(offs, lits, model)


@pytest.mark.parametrize("name", sorted(k for k in G["verdicts"] if "as-shipped" not in k))
def test_verdicts_match_reference_golden(name):
    parts = name.split("-")
    if parts[0].startswith("uf"):
        offs, lits = random_ksat(int(parts[0][2:]), int(parts[1]), int(parts[2][4:]))
    else:
        offs, lits = pigeonhole(int(parts[1]), int(parts[2]))
    verdict, model, cnf, pre = oracle_verdict(offs, lits)
    assert verdict == G["verdicts"][name]
    if verdict == g.SAT:
        assert check_model(offs, lits, model)


def test_iteration_cap_reproduces_reference_undef():
    # the shipped reference answers UNDEFINED on uf100 (MAX_ITERATIONS 1000, SATSolver/Configs.cuh:23); with the
    # reference decision rule and the same cap the oracle also runs out of iterations
    assert G["verdicts"]["uf100-426-seed0-as-shipped"] == g.UNDEF
    offs, lits = random_ksat(100, 426, 0)
    verdict, _, _, _ = oracle_verdict(offs, lits, decision=0, max_iterations=1000)
    assert verdict == g.UNDEF


@pytest.mark.parametrize("seed", [0, 1])
def test_bcp_sets_match_reference_golden(seed):
    d = G["cubes"][f"uf250-1065-seed{seed}"]
    offs, lits = random_ksat(250, 1065, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cubes = pre.choose_cubes(8, 32)
    o = Oracle(250, pre.offsets, pre.lits)
    r = o.run(*cube_csr(cubes), mode=1)
    for p in d["propagate"]:
        j = p["cube"]
        assert r["records"]["status"][j] == p["status"]
        if p["status"] == g.UNDEF:
            assert sorted(r["implied"][j, : r["n_implied"][j]].tolist()) == p["implied_sorted"]
        else:
            c = r["conflict_clause"][j]        # reported clause is falsified under cube + implied literals
            assert c >= 0
            trail = set(cubes[j].tolist()) | set(r["implied"][j, : r["n_implied"][j]].tolist())
            assert all((x ^ 1) in trail for x in pre.lits[pre.offsets[c]: pre.offsets[c + 1]])


@pytest.mark.skipif(not have_ref(), reason="reference host build not present")
def test_bcp_sets_match_reference_live_all_cubes(quiet):
    from oracle.binding import Reference
    offs, lits = random_ksat(250, 1065, 2)
    with quiet():
        R = Reference(offs, lits)
        poff, plits = R.formula()
        cubes = R.cubes(8, 32, 0)
    o = Oracle(250, poff, plits)
    r = o.run(*cube_csr(cubes), mode=1)
    bad = 0
    with quiet():
        for j in range(len(cubes)):
            st, imp = R.propagate(cubes[j])
            if st != r["records"]["status"][j]:
                bad += 1
            elif st == g.UNDEF and set(imp.tolist()) != set(r["implied"][j, : r["n_implied"][j]].tolist()):
                bad += 1
    assert bad == 0


@pytest.mark.skipif(not have_ref(), reason="reference host build not present")
@pytest.mark.parametrize("n,m", [(20, 91), (30, 128), (50, 218)])
def test_verdicts_match_reference_live(n, m, quiet):
    from oracle.binding import Reference
    for seed in range(20, 26):
        offs, lits = random_ksat(n, m, seed)
        with quiet():
            R = Reference(offs, lits)
            st, model = R.solve() if R.status == 2 else (R.status, [])
        verdict, mine, cnf, pre = oracle_verdict(offs, lits)
        assert verdict == st
        if st == g.SAT and R.status == 2:
            poff, plits = pre.offsets, pre.lits
            assert check_model(poff, plits, model_from_lits(cnf.n_vars, model))
            assert check_model(offs, lits, mine)


def test_learnt_clauses_are_implied_and_asserting():
    """every learnt clause must be implied by the formula.  Cheap sound check: solving
    formula + negated learnt clause cube is UNSAT is expensive; instead check the run invariants the oracle exposes:
    a cube run and a no-cube run agree on the verdict, and UNSAT cubes stay UNSAT with the other decision rule."""
    offs, lits = random_ksat(60, 258, 4)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    o = Oracle(60, pre.offsets, pre.lits)
    whole = o.run(*EMPTY)["records"]["status"][0]
    cubes = pre.choose_cubes(1, 4)
    a = o.run(*cube_csr(cubes), stop_on_sat=False, decision=1)["records"]["status"]
    b = o.run(*cube_csr(cubes), stop_on_sat=False, decision=0)["records"]["status"]
    assert np.array_equal(a, b)
    assert (whole == g.SAT) == bool((a == g.SAT).any())


def test_eval_clauses_semantics():
    # VariablesStateHandler::clause_status: SAT wins immediately; unit literal = LAST unassigned literal
    offs = np.array([0, 3, 6, 8], dtype=np.int64)
    lits = np.array([1, 3, 5, 0, 2, 4, 7, 6], dtype=np.int32)       # (x0 x1 x2) (~x0 ~x1 ~x2) (x3 ~x3)
    o = Oracle(4, offs, lits)
    T, F, U = 0, 1, 2                                               # reference sat_status encoding
    st, unit = o.eval_clauses(np.array([[F, F, U, U], [T, T, T, U], [U, U, U, U], [F, F, F, T]], dtype=np.uint8))
    assert st.tolist() == [[2, 0, 2], [0, 1, 2], [2, 2, 2], [1, 0, 0]]
    assert unit.tolist() == [[5, -1, -1], [-1, -1, -1], [-1, -1, -1], [-1, -1, -1]]


# ---- clause evaluation pinned by EXECUTING the reference's VariablesStateHandler::clause_status
# (SATSolver/VariablesStateHandler.cu:180-206) — live where oracle/_ref is built from /root/reference, and through the
# committed fixture (tests/golden/reference_outputs.json: "clause_status") everywhere else
@pytest.mark.parametrize("name,n,m,seed", [("uf50-218-seed0", 50, 218, 0), ("uf250-1065-seed0", 250, 1065, 0)])
def test_eval_clauses_equals_reference_golden(name, n, m, seed):
    offs, lits = random_ksat(n, m, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    o = Oracle(n, pre.offsets, pre.lits)
    for row in G["clause_status"][name]:
        a = np.array(row["assignment"], dtype=np.uint8)
        st, unit = o.eval_clauses(a[None, :])
        assert st[0].tolist() == row["status"]
        assert unit[0].tolist() == row["unit"]


def test_eval_clauses_equals_reference_live(reference_available, quiet):
    if not reference_available:
        pytest.skip("reference build not present")
    offs, lits = random_ksat(100, 426, 5)
    with quiet():
        R = Reference(offs, lits)
        poff, plits = R.formula()
    o = Oracle(100, poff, plits)
    rg = np.random.default_rng(99)
    for frac in (0.1, 0.5, 0.9, 1.0):
        a = np.full(100, 2, dtype=np.uint8)
        pick = rg.random(100) < frac
        a[pick] = rg.integers(0, 2, size=int(pick.sum()), dtype=np.uint8)
        with quiet():
            st_ref, unit_ref = R.clause_status(a)
        st, unit = o.eval_clauses(a[None, :])
        assert st[0].tolist() == st_ref.tolist() and unit[0].tolist() == unit_ref.tolist()
