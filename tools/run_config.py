#!/usr/bin/env python
"""Runs one BASELINE.json-style configuration on one GPU through the C ABI and prints a JSON line.
usage: python tools/run_config.py php P H | uf N M SEED   [--bt B T] [--time-limit S] [--share L] [--epoch-ms E]
The solve runs as budgeted steps so that a wall-clock limit can end it (verdict UNDEF, jobs_done tells how far it got)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpupsat_b200 as g
from gpupsat_b200.instances import check_model, pigeonhole, random_ksat

ap = argparse.ArgumentParser()
ap.add_argument("kind")
ap.add_argument("args", nargs="+", type=int)
ap.add_argument("--bt", nargs=2, type=int, default=[8, 32])
ap.add_argument("--time-limit", type=float, default=60.0)
ap.add_argument("--share", type=int, default=0)
ap.add_argument("--epoch-ms", type=float, default=50.0)
a = ap.parse_args()
if a.kind == "php":
    offs, lits = pigeonhole(*a.args)
    name = f"PHP({a.args[0]},{a.args[1]})"
else:
    n, m, seed = a.args
    offs, lits = random_ksat(n, m, seed)
    name = f"uf{n}-{m} seed {seed}"
cnf = g.Cnf.from_arrays(offs, lits)
pre = cnf.preprocess()
cubes = pre.choose_cubes(*a.bt)
opts = dict(share_learnts=1, share_max_len=a.share) if a.share else {}
with g.Solver(cnf.n_vars, pre.offsets, pre.lits, **opts) as s:
    s.set_cubes(cubes)
    t0 = time.time()
    s.solve_begin()
    steps = 0
    while True:
        done, verdict = s.solve_step(budget_ms=a.epoch_ms)
        steps += 1
        if done or time.time() - t0 > a.time_limit:
            break
    verdict, model, st = s.solve_end()
    wall = time.time() - t0
ok = None
if verdict == g.SAT:
    ok = bool(check_model(pre.offsets, pre.lits, model))
print(json.dumps({"instance": name, "n_vars": cnf.n_vars, "clauses": len(pre.offsets) - 1, "cubes": len(cubes),
                  "cube_literals": int(cubes.shape[1]), "verdict": {0: "SAT", 1: "UNSAT", 2: "UNDEF"}[verdict],
                  "model_verified": ok, "kernel_ms": st["kernel_ms"], "wall_s": wall, "steps": steps,
                  "jobs_done": st["jobs_done"], "implications": st["implications"], "conflicts": st["conflicts"],
                  "implications_per_s": st["implications"] / (st["kernel_ms"] * 1e-3),
                  "conflicts_per_s": st["conflicts"] / (st["kernel_ms"] * 1e-3), "splits": st["splits"],
                  "warp_busy_frac": st["warp_busy_frac"], "blocks": st["blocks"], "warps_per_block": st["warps_per_block"],
                  "state_in_smem": st["state_in_smem"]}))
