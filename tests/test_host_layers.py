"""Host layers behind the C ABI (no GPU): DIMACS reader, preprocessing, cube generation — against the golden
outputs of the reference's own code (tests/golden/reference_outputs.json) and, where it is built, the reference."""
import os

import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import parse_dimacs_text, pigeonhole, random_ksat, to_dimacs
from oracle.binding import Reference
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


# ---- index of the large-database sweep kernel (host_formula.cpp: build_sweep_index / order_cubes_for_sweep) ----------
def _decode_bucket(words):
    """(count field, list-pointer field, [(a, b)] * 11) of one 64-byte bucket, read the way the kernel does (tern_field
    in kernels.cu)"""
    bits = 0
    for i, w in enumerate(words):
        bits |= int(w) << (32 * i)
    entries = []
    for j in range(11):
        o = 32 + 42 * j if j < 5 else 256 + 42 * (j - 5)
        entries.append(((bits >> o) & 0x1FFFFF, (bits >> (o + 21)) & 0x1FFFFF))
    return int(words[0]) & 255, int(words[0]) >> 8, entries


def test_bucket_index_holds_every_occurrence_list():
    from gpupsat_b200.instances import random_ksat
    from tests.emu import binding as emu
    n, m = 300, 1300
    offs, lits = random_ksat(n, m, 3)
    extra = [[1, 2 * v, 2 * v + 3] for v in range(2, 32)]                     # literal 1 gets a list of 30+ entries
    extra += [[3, 2 * v, 2 * v + 5] for v in range(3, 290)]                    # literal 3 more than the count field holds
    lits = np.concatenate([lits, np.array(extra, dtype=np.int32).reshape(-1)])
    offs = np.arange(0, len(lits) + 1, 3, dtype=np.int64)
    bucket, orange = emu.bucket_index(n, offs, lits)
    clauses = lits.reshape(-1, 3)
    want = {f: [] for f in range(2 * n)}
    for c in clauses:                                                          # clause order = occurrence order
        for i in range(3):
            want[int(c[i])].append(tuple(int(x) for k, x in enumerate(c) if k != i))
    pad = 2 * n + 1
    assert max(len(v) for v in want.values()) > 11
    assert max(len(v) for v in want.values()) > 255
    for f in range(2 * n):
        cnt8, ptr, entries = _decode_bucket(bucket[f])
        cnt = len(want[f])
        assert cnt8 == min(cnt, 255) and orange[f, 1] - orange[f, 0] in (cnt, cnt + 1)
        assert 2 * ptr == orange[f, 0]                                         # where entries 11.. are read from
        for j in range(11):
            assert entries[j] == (want[f][j] if j < cnt else (pad, pad)), (f, j)
    for f in (2 * n, 2 * n + 1):                                               # the sentinel's buckets: empty, all padding
        cnt, ptr, entries = _decode_bucket(bucket[f])
        assert cnt == 0 and all(e == (pad, pad) for e in entries)
    # not pure 3-SAT -> no bucket index
    assert emu.bucket_index(4, np.array([0, 2, 5]), np.array([0, 3, 1, 4, 6], dtype=np.int32)) is None


def test_cube_order_for_the_sweep_kernel():
    from gpupsat_b200.instances import random_ksat
    from tests.emu import binding as emu
    n, m = 300, 1300
    offs, lits = random_ksat(n, m, 4)
    occ = np.bincount(lits, minlength=2 * n)
    rng = np.random.default_rng(0)
    cubes = [rng.choice(2 * n, size=k, replace=False).astype(np.int32) for k in (40, 1, 0, 77)]
    cubes[1] = np.array([5], dtype=np.int32)
    cubes.append(np.array([8, 9, 8, 20], dtype=np.int32))                      # variable 4 three times
    co = np.cumsum([0] + [len(c) for c in cubes]).astype(np.int64) + 3         # a base offset, as the ABI allows
    cl = np.concatenate([np.zeros(3, dtype=np.int32)] + cubes)
    out, info = emu.order_cubes(n, offs, lits, co, cl)
    at = 0
    for j, c in enumerate(cubes):
        got = out[at: at + len(c)]
        at += len(c)
        assert sorted(got.tolist()) == sorted(c.tolist())                      # a permutation of the cube
        cls = np.minimum(occ[got ^ 1], 11)
        assert (np.diff(cls) >= 0).all()                                       # ordered by the list length of the negation
        for k in range(12):                                                    # stable inside a class
            assert got[cls == k].tolist() == [x for x in c.tolist() if min(occ[x ^ 1], 11) == k]
        assert (info[j] & 0x3FFFFFFF) == int((occ[c ^ 1] <= 5).sum())
        distinct = len(set((c >> 1).tolist())) == len(c)
        assert ((info[j] >> 30) & 1) == int(distinct)
    # random_ksat rejects repeats inside a clause but a random 40-literal cube may hold x and ~x: both cases occurred?
    assert ((info >> 30) & 1).tolist()[4] == 0


def test_ternary_state_arithmetic_of_the_sweep_kernel():
    """the integer identities gpsat_bcp_sweep_tern_kernel's lookups rest on (kernels.cu: TernJob::code / assign)"""
    x = np.arange(0, 1 << 21, dtype=np.uint64)
    assert np.array_equal((x * 0x1999999A) >> 32, x // 10)                     # q = __umulhi(x, 0x1999999A) = x / 10
    r = np.arange(5)
    assert [(0x1B090301 >> (8 * int(k))) & 255 if k < 4 else 81 for k in r] == [1, 3, 9, 27, 81]
    # five base-3 digits per byte never exceed a byte, and adding a digit's weight to a byte whose digit is 0 changes
    # that digit only (what the add-assigned cubes rely on)
    assert 2 * (1 + 3 + 9 + 27 + 81) == 242
    for byte in range(243):
        digits = [(byte // 3 ** k) % 3 for k in range(5)]
        for k in range(5):
            if digits[k] == 0:
                for d in (1, 2):
                    nb = byte + d * 3 ** k
                    assert nb <= 242 and [(nb // 3 ** i) % 3 for i in range(5)] == digits[:k] + [d] + digits[k + 1:]
    # literal x = 2 var + sign lives in byte x // 10 at table column x % 10 = 2 (var % 5) + sign
    v = np.arange(0, 1 << 20, dtype=np.int64)
    for sgn in (0, 1):
        lit = 2 * v + sgn
        assert np.array_equal(lit // 10, v // 5) and np.array_equal(lit % 10, 2 * (v % 5) + sgn)


# ---- SimpleJobChooser (JobsManager/SimpleJobChooser.cu:22-75, behind USE_SIMPLE_JOBS_GENERATION) = GPSAT_STRATEGY_SIMPLE
@pytest.mark.parametrize("name", ["preprocess-case3", "uf250-1065-seed0"])
def test_simple_job_chooser_matches_reference_golden(name):
    d = G["cubes"]["simple"][name]
    if name == "uf250-1065-seed0":
        offs, lits = random_ksat(250, 1065, 0)
    else:
        pc = G["preprocess"]["case3"]
        offs, lits = np.array(pc["offsets"], dtype=np.int64), np.array(pc["lits"], dtype=np.int32)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cs = pre.choose_cubes(8, 32, g.binding.STRATEGY_SIMPLE)
    assert cs.shape == (d["n"], d["k"]) and (cs[0] >> 1).tolist() == d["vars"]
    assert cs[1].tolist() == d["second"] and cs[-1].tolist() == d["last"] and cube_checksum(cs) == d["checksum"]
    assert not set((cs[0] >> 1).tolist()) & set((pre.solved >> 1).tolist())      # solved (dead) variables are skipped


def test_simple_job_chooser_matches_reference_live(reference_available, quiet):
    if not reference_available:
        pytest.skip("reference build not present")
    for seed in (3, 4):
        offs, lits = random_ksat(40, 170, seed)
        with quiet():
            want = Reference(offs, lits).cubes_simple()
        got = g.Cnf.from_arrays(offs, lits).preprocess().choose_cubes(1, 1, g.binding.STRATEGY_SIMPLE)
        assert np.array_equal(got, want)


# ---- launch geometry of the CDCL kernel (gpsat_device.h: gpsat_plan_warps) ----------------------------------------------
def test_cdcl_launch_geometry_is_pinned():
    """the arithmetic that decides how many warps of job state share an SM with the staged formula (B200: 227 KB
    opt-in); the measured optimum for config 2 is 24 warps x 6.8 KB beside the 31 KB packed formula"""
    from tests.emu import binding as emu
    c2 = emu.plan(250, 3195, 1065)                                        # uf250-1065
    assert c2 == {"warps": 24, "state_in_smem": 1, "formula_in_smem": 1, "formula_smem_words": 7960, "smem_bytes": 194656,
                  "state_words": 1696, "idx16": 1, "lbuf_words": 251}
    assert emu.plan(250, 3195, 1065, phase_stats=1)["state_words"] == 1696 + 28   # the statistic words only when kept
    asked = emu.plan(250, 3195, 1065, warps=28)
    assert asked["warps"] == 28 and asked["formula_in_smem"] == 1 and asked["smem_bytes"] <= 232448
    assert emu.plan(250, 3195, 1065, warps=64)["warps"] == 28             # clamped to the widest kernel variant
    # propagate-only runs stage nothing and keep 32-bit state
    bcp = emu.plan(250, 3195, 1065, solve=0)
    assert bcp["formula_in_smem"] == 0 and bcp["idx16"] == 0 and bcp["state_in_smem"] == 1 and bcp["warps"] == 24
    # a formula whose slots do not fit 16 bits is not staged (and its state stays 32-bit) ...
    mid = emu.plan(1500, 66000, 22000)
    assert mid["formula_in_smem"] == 0 and mid["idx16"] == 0 and mid["state_in_smem"] == 1 and mid["warps"] == 4
    assert mid["smem_bytes"] == mid["warps"] * mid["state_words"] * 4 <= 232448
    # ... one that fits 16 bits but would leave no room for 8 warps is not staged either
    tight = emu.plan(1000, 12780, 4260)
    assert tight["formula_in_smem"] == 0 and tight["idx16"] == 0 and tight["state_in_smem"] == 1 and tight["warps"] == 7
    # ... one that leaves room for 9 is
    roomy = emu.plan(600, 7668, 2556)
    assert roomy["formula_in_smem"] == 1 and roomy["idx16"] == 1 and roomy["warps"] == 9
    assert roomy["smem_bytes"] == 4 * (roomy["formula_smem_words"] + 9 * roomy["state_words"]) <= 232448
    # state too large for 4 warps (n = 5000), a million variables: state in global memory, 16 warps
    for big in (emu.plan(5000, 63900, 21300), emu.plan(1_000_000, 12_000_000, 4_000_000)):
        assert (big["warps"], big["state_in_smem"], big["formula_in_smem"], big["smem_bytes"], big["idx16"]) == (16, 0, 0, 0, 0)
    # a requested geometry that does not fit falls back to global state instead of failing
    forced = emu.plan(1500, 66000, 22000, warps=24)
    assert forced["state_in_smem"] == 0 and forced["warps"] == 24 and forced["smem_bytes"] == 0
    # PHP(10,9): small state, the warp count is the cap, not the memory
    php = emu.plan(90, 1900, 415)
    assert php["warps"] == 24 and php["formula_in_smem"] == 1
