#!/usr/bin/env python
"""Dev helper: per-phase warp time of one C2 solve (opts.phase_stats), as shares of the summed job time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat
offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)
with g.Solver(250, pre.offsets, pre.lits, phase_stats=1) as s:
    s.set_cubes(cubes)
    for _ in range(3):
        v, m, st = s.solve()
    ph = s.phase_stats()
warps = st["blocks"] * st["warps_per_block"]
print("kernel ms", round(st["kernel_ms"], 2), "warps", warps, "jobs", ph["jobs"], "job_ns sum / (warps x kernel)", round(ph["job_ns"] / (warps * st["kernel_ms"] * 1e6), 3),
      "idle share", round(ph["idle_ns"] / (warps * st["kernel_ms"] * 1e6), 3))
for i in range(8):
    print(i, "ns share of job time", round(ph["ns"][i] / max(ph["job_ns"], 1), 4), "count", int(ph["count"][i]), "mean us", round(ph["ns"][i] / max(ph["count"][i], 1) / 1e3, 2))
print("unaccounted share of job time", round(1 - ph["ns"].sum() / max(ph["job_ns"], 1), 4), "mean job us", round(ph["job_ns"] / max(ph["jobs"], 1) / 1e3, 1))
