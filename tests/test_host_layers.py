"""Host layers behind the C ABI (no GPU): DIMACS reader, preprocessing, cube generation — against the golden
outputs of the reference's own code (tests/golden/reference_outputs.json) and, where it is built, the reference."""
import os

import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import parse_dimacs_text, pigeonhole, random_ksat, to_dimacs
from tests.helpers import cube_checksum, golden, have_ref

G = golden()


@pytest.mark.parametrize("name", sorted(G["tests_cnf"]))
def test_dimacs_reader_on_reference_fixtures(tmp_path, name):
    text = G["tests_cnf"][name]["dimacs"]
    p = tmp_path / name
    p.write_text(text)
    cnf = g.Cnf.read(str(p))
    offs, lits, nv = parse_dimacs_text(text)
    assert np.array_equal(cnf.offsets, offs)
    assert np.array_equal(cnf.lits, lits)
    assert cnf.n_vars == int(lits.max() >> 1) + 1
    assert cnf.header_vars == nv
    assert cnf.n_lines == text.count("\n") + (1 if text.count("\n") else 0)
    pre = cnf.preprocess()
    assert pre.status == G["tests_cnf"][name]["pre_status"]
    assert pre.solved.tolist() == G["tests_cnf"][name]["solved"]


def test_dimacs_reader_quirks(tmp_path):
    # comments only before the header; clause block ends at a blank line; clause lines must close with 0
    ok = tmp_path / "ok.cnf"
    ok.write_text("c hello\n\nc again\np cnf 3 2\n1 -2 0\n 2  3 0 \n\nthis is ignored\n")
    cnf = g.Cnf.read(str(ok))
    assert cnf.n_clauses == 2 and cnf.lits.tolist() == [1, 2, 3, 5]
    empty_clause = tmp_path / "e.cnf"
    empty_clause.write_text("p cnf 2 2\n1 2 0\n0\n")
    assert g.Cnf.read(str(empty_clause)).offsets.tolist() == [0, 2, 2]
    for bad in ("p cnf 2 1\n1 2\n", "1 2 0\n", "p cnf 2 1\n1 2 0\nc late comment\n", "p cnf 2 1\n1 x 0\n",
                "p dnf 2 1\n1 2 0\n"):
        f = tmp_path / "bad.cnf"
        f.write_text(bad)
        with pytest.raises(g.GpsatError) as e:
            g.Cnf.read(str(f))
        assert e.value.code == -5
    with pytest.raises(g.GpsatError) as e:
        g.Cnf.read(str(tmp_path / "missing.cnf"))
    assert e.value.code == -4
    # header not trusted: n_vars = highest variable seen
    h = tmp_path / "h.cnf"
    h.write_text("p cnf 2 7\n1 -9 0\n")
    c = g.Cnf.read(str(h))
    assert c.n_vars == 9 and c.header_vars == 2 and c.header_clauses == 7 and c.n_clauses == 1


@pytest.mark.parametrize("case", sorted(G["preprocess"]))
def test_preprocess_matches_reference_golden(case):
    d = G["preprocess"][case]
    cnf = g.Cnf.from_arrays(np.array(d["offsets"]), np.array(d["lits"]))
    pre = cnf.preprocess()
    assert cnf.n_vars == d["n_vars"]
    assert pre.status == d["status"]
    assert pre.solved.tolist() == d["solved"]
    if d["status"] == g.UNDEF:
        assert pre.offsets.tolist() == d["pre_offsets"]
        assert pre.lits.tolist() == d["pre_lits"]


@pytest.mark.skipif(not have_ref(), reason="reference host build not present")
@pytest.mark.parametrize("seed", range(20))
def test_preprocess_matches_reference_live(seed, quiet):
    from oracle.binding import Reference
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 40))
    m = int(rng.integers(5, 120))
    cl = []
    for _ in range(m):
        ln = int(rng.choice([1, 1, 2, 2, 3, 3, 3, 4]))
        cl.append([int(x) for x in rng.integers(0, 2 * n, size=ln)])     # repeats and tautologies on purpose
    offs = np.cumsum([0] + [len(c) for c in cl]).astype(np.int64)
    lits = np.array([x for c in cl for x in c], dtype=np.int32)
    with quiet():
        R = Reference(offs, lits)
        poff, plits = R.formula()
        solved = R.solved_literals()
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    assert pre.status == R.status
    assert pre.solved.tolist() == solved.tolist()
    if R.status == g.UNDEF:
        assert pre.offsets.tolist() == poff.tolist()
        assert pre.lits.tolist() == plits.tolist()


@pytest.mark.parametrize("name", ["uf250-1065-seed0", "uf250-1065-seed1", "php-10-9"])
def test_cubes_match_reference_golden(name):
    d = G["cubes"][name]
    if name.startswith("uf250"):
        offs, lits = random_ksat(250, 1065, int(name[-1]))
    else:
        offs, lits = pigeonhole(10, 9)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cb = pre.choose_cubes(8, 32, g.STRATEGY_DISTRIBUTED)
    assert cb.shape == (d["n"], d["k"])
    assert cb[0].tolist() == d["first"] and cb[1].tolist() == d["second"] and cb[-1].tolist() == d["last"]
    assert cube_checksum(cb) == d["checksum"]
    if "uniform_k" in d:
        cu = pre.choose_cubes(8, 32, g.STRATEGY_UNIFORM)
        assert cu.shape[1] == d["uniform_k"] and (cu[0] >> 1).tolist() == d["uniform_vars"]
    # cube j, position i positive iff bit (k-1-i) of j is 0
    k = d["k"]
    j = 0b101100111010 % d["n"]
    want = [2 * v + (0 if (j >> (k - 1 - i)) & 1 else 1) for i, v in enumerate(d["vars"])]
    assert cb[j].tolist() == want


def test_cube_count_caps():
    d = G["cubes"]["sat_8v_random-b2-t2"]
    offs, lits, _ = parse_dimacs_text(G["tests_cnf"]["sat_8v_random.cnf"]["dimacs"])
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cb = pre.choose_cubes(2, 2)
    assert cb.shape == (d["n"], d["k"]) and (cb[0] >> 1).tolist() == d["vars"]
    # MAX_VARS 15 cap
    offs, lits = random_ksat(250, 1065, 3)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    assert pre.choose_cubes(1024, 1024).shape == (1 << 15, 15)
    assert pre.choose_cubes(1, 1).shape[1] == 4          # 10 jobs wanted -> floor(log2 10) + 1


def test_roundtrip_dimacs(tmp_path):
    offs, lits = random_ksat(40, 170, 9)
    p = tmp_path / "r.cnf"
    p.write_text(to_dimacs(offs, lits, 40))
    cnf = g.Cnf.read(str(p))
    assert np.array_equal(cnf.offsets, offs) and np.array_equal(cnf.lits, lits)
    assert cnf.largest_clause == 3
    v, cnt = np.unique(lits >> 1, return_counts=True)
    assert cnf.most_common_freq == cnt.max()
