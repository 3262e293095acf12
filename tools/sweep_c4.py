#!/usr/bin/env python
"""Config 4 sweep: HBM-bound BCP throughput on a large clause database (planted 3-SAT n=1e6, m=4e6).
For J jobs x L-literal trails: literals propagated / s, implications / s, counted algorithmic bytes / s vs HBM peak.
usage: python tools/sweep_c4.py [--jobs 148,1184,4736] [--lens 50000,100000,200000] [--out profiles/r01_c4_sweep.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpupsat_b200 as g  # noqa: E402
from gpupsat_b200.instances import planted_3sat_large, sweep_trails  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", default="148,1184,4736")
    ap.add_argument("--lens", default="50000,100000,200000")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--flags", default="0", help="gpsat_opts.sweep_flags, comma list (test hooks: 2 register prefetch, 4 no L2 prefetch, "
                    "32 lane-private code table, 64 first hit kept during the scan)")
    args = ap.parse_args()
    n, m = 1_000_000, 4_000_000
    offs, lits, planted = planted_3sat_large(n, m, 4)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    rows = []
    for flags in [int(x) for x in args.flags.split(",")]:
      with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE, sweep_flags=flags) as s:
        for L in [int(x) for x in args.lens.split(",")]:
            for J in [int(x) for x in args.jobs.split(",")]:
                co, cl = sweep_trails(n, J, L, 4, planted)
                stride = min(n, 6 * L)
                s.set_cubes(cube_offsets=co, cube_lits=cl)
                best = None
                for r in range(args.reps + 1):
                    t = time.perf_counter()
                    got = s.propagate_all(implied_stride=stride, want_implied=False)
                    rec = got["records"]
                    kms = s.last_kernel_ms()          # CUDA-event time of the sweep kernel
                    if r > 0 and (best is None or kms < best):
                        best = kms
                imp = int(got["n_implied"].sum())
                visited = int(rec["watchers_visited"].sum())
                words = int(rec["clause_words_read"].sum())
                props = J * L + imp                      # trail literals whose occurrence lists were walked
                alg_bytes = 8 * visited + 4 * words + 8 * imp
                row = {"sweep_flags": flags, "jobs": J, "trail_len": L, "kernel_ms": best, "implications": imp, "literals_propagated": props,
                       "entries_visited": visited, "implications_per_s": imp / (best * 1e-3),
                       "literals_per_s": props / (best * 1e-3), "algorithmic_GBps": alg_bytes / (best * 1e-3) / 1e9,
                       "hbm_frac": alg_bytes / (best * 1e-3) / 1e9 / peak, "status_undef": int((rec["status"] == 2).sum())}
                rows.append(row)
                print(json.dumps(row), flush=True)
    if args.out:
        json.dump({"instance": "planted 3-SAT n=1e6 m=4e6 seed 4", "hbm_peak_GBps": peak, "rows": rows},
                  open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
