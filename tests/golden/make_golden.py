"""Generates tests/golden/*.json by running the REFERENCE's own code (oracle/_ref/libgpsat_ref.so, built from
/root/reference/src by oracle/ref/build_ref.sh) in the build container.  /root/reference does not exist on the GPU
box, so these fixtures are what pins the oracle and the product there.  Re-run: python tests/golden/make_golden.py"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gpupsat_b200.instances import parse_dimacs_text, pigeonhole, random_ksat  # noqa: E402
from oracle.binding import Quiet, Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_solve_full(offs, lits, nocap=False):
    with Quiet():
        R = Reference(offs, lits, nocap=nocap)
        if R.status != 2:
            return {"pre_status": int(R.status), "verdict": int(R.status), "solved": R.solved_literals().tolist()}
        st, model = R.solve()
    return {"pre_status": 2, "verdict": int(st), "solved": R.solved_literals().tolist(), "model_lits": model.tolist()}


def main():
    out = {}
    # 1. the reference's own tests/cnf (verdict + .expected) — the files themselves are tiny; embed their text
    cnf_dir = "/root/reference/tests/cnf"
    files = {}
    for f in sorted(glob.glob(os.path.join(cnf_dir, "*.cnf"))):
        text = open(f).read()
        offs, lits, nv = parse_dimacs_text(text)
        r = ref_solve_full(offs, lits)
        r["expected"] = open(f.replace(".cnf", ".expected")).read().strip()
        r["dimacs"] = text
        files[os.path.basename(f)] = r
    out["tests_cnf"] = files

    # 2. verdicts of the reference (as shipped) on generated instances
    verdicts = {}
    for n, m in ((20, 91), (50, 218)):
        for seed in range(10):
            offs, lits = random_ksat(n, m, seed)
            verdicts[f"uf{n}-{m}-seed{seed}"] = ref_solve_full(offs, lits)["verdict"]
    for p, h in ((4, 3), (5, 4), (6, 5)):
        offs, lits = pigeonhole(p, h)
        verdicts[f"php-{p}-{h}"] = ref_solve_full(offs, lits)["verdict"]
    offs, lits = random_ksat(100, 426, 0)
    verdicts["uf100-426-seed0-nocap"] = ref_solve_full(offs, lits, nocap=True)["verdict"]
    verdicts["uf100-426-seed0-as-shipped"] = ref_solve_full(offs, lits)["verdict"]
    out["verdicts"] = verdicts

    # 3. preprocessing: formula after the reference's host preprocessing on instances with unit clauses / repeats
    pre = {}
    rng = np.random.default_rng(7)
    for case in range(12):
        n, m = 30, 70
        offs, lits = random_ksat(n, m, 100 + case)
        cl = [lits[offs[i]:offs[i + 1]].tolist() for i in range(m)]
        for _ in range(3):   # units
            cl.insert(int(rng.integers(0, len(cl))), [int(rng.integers(0, 2 * n))])
        cl.insert(int(rng.integers(0, len(cl))), [4, 4, 9])          # repeated literal
        cl.insert(int(rng.integers(0, len(cl))), [6, 7, 11])         # tautology
        cl.insert(int(rng.integers(0, len(cl))), [cl[0][0], cl[0][0]])   # collapses to a unit after repeat removal
        o2 = np.cumsum([0] + [len(c) for c in cl]).astype(np.int64)
        l2 = np.array([x for c in cl for x in c], dtype=np.int32)
        with Quiet():
            R = Reference(o2, l2)
            poff, plits = R.formula()
            solved = R.solved_literals()
        pre[f"case{case}"] = {"offsets": o2.tolist(), "lits": l2.tolist(), "status": int(R.status),
                              "n_vars": int(R.n_vars), "pre_offsets": poff.tolist(), "pre_lits": plits.tolist(),
                              "solved": solved.tolist()}
    out["preprocess"] = pre

    # 4. cubes (MaxClauseJobChooser) and BCP of cubes (set_assumptions): uf250 seeds 0,1 with -b 8 -t 32
    cubes = {}
    for seed in (0, 1):
        offs, lits = random_ksat(250, 1065, seed)
        with Quiet():
            R = Reference(offs, lits)
            cb = R.cubes(8, 32, 0)
            cu = R.cubes(8, 32, 1)
            props = []
            for j in range(0, len(cb), 16):          # every 16th cube: status + implied SET
                st, imp = R.propagate(cb[j])
                props.append({"cube": j, "status": int(st), "implied_sorted": sorted(imp.tolist())})
        cubes[f"uf250-1065-seed{seed}"] = {"k": int(cb.shape[1]), "n": int(len(cb)), "vars": (cb[0] >> 1).tolist(),
                                          "first": cb[0].tolist(), "second": cb[1].tolist(), "last": cb[-1].tolist(),
                                          "checksum": int((cb.astype(np.int64) * (np.arange(cb.size).reshape(cb.shape) % 1009 + 1)).sum()),
                                          "uniform_k": int(cu.shape[1]), "uniform_vars": (cu[0] >> 1).tolist(),
                                          "propagate": props}
    for name, (offs, lits) in {"php-10-9": pigeonhole(10, 9)}.items():
        with Quiet():
            R = Reference(offs, lits)
            cb = R.cubes(8, 32, 0)
        cubes[name] = {"k": int(cb.shape[1]), "n": int(len(cb)), "vars": (cb[0] >> 1).tolist(), "first": cb[0].tolist(),
                       "second": cb[1].tolist(), "last": cb[-1].tolist(),
                       "checksum": int((cb.astype(np.int64) * (np.arange(cb.size).reshape(cb.shape) % 1009 + 1)).sum())}
    # small-formula k caps (n_live - 2) from tests/cnf
    text = open(os.path.join(cnf_dir, "sat_8v_random.cnf")).read()
    offs, lits, _ = parse_dimacs_text(text)
    with Quiet():
        R = Reference(offs, lits)
        cb = R.cubes(2, 2, 0)
    cubes["sat_8v_random-b2-t2"] = {"k": int(cb.shape[1]), "n": int(len(cb)), "vars": (cb[0] >> 1).tolist()}
    # SimpleJobChooser (USE_SIMPLE_JOBS_GENERATION): on a formula with solved (dead) variables and on uf250
    simple = {}
    pc = out["preprocess"]["case3"]
    for name, (o2, l2) in {"preprocess-case3": (np.array(pc["offsets"], dtype=np.int64), np.array(pc["lits"], dtype=np.int32)),
                           "uf250-1065-seed0": random_ksat(250, 1065, 0)}.items():
        with Quiet():
            R = Reference(o2, l2)
            cs = R.cubes_simple()
        simple[name] = {"k": int(cs.shape[1]), "n": int(len(cs)), "vars": (cs[0] >> 1).tolist(), "second": cs[1].tolist(),
                        "last": cs[-1].tolist(),
                        "checksum": int((cs.astype(np.int64) * (np.arange(cs.size).reshape(cs.shape) % 1009 + 1)).sum())}
    cubes["simple"] = simple
    out["cubes"] = cubes

    # 5. clause evaluation: VariablesStateHandler::clause_status of every clause under seeded partial assignments
    ev = {}
    for name, (n, m, seed) in {"uf50-218-seed0": (50, 218, 0), "uf250-1065-seed0": (250, 1065, 0)}.items():
        offs, lits = random_ksat(n, m, seed)
        rows = []
        rg = np.random.default_rng(1234 + seed + n)
        with Quiet():
            R = Reference(offs, lits)
            for frac in (0.0, 0.3, 0.7, 0.95, 1.0):
                a = np.full(n, 2, dtype=np.uint8)
                pick = rg.random(n) < frac
                a[pick] = rg.integers(0, 2, size=int(pick.sum()), dtype=np.uint8)
                st, unit = R.clause_status(a)
                rows.append({"assignment": a.tolist(), "status": st.tolist(), "unit": unit.tolist()})
        ev[name] = rows
    out["clause_status"] = ev

    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(out, f, indent=None, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "reference_outputs.json"), os.path.getsize(os.path.join(HERE, "reference_outputs.json")), "bytes")


if __name__ == "__main__":
    main()
