#!/usr/bin/env python
"""Dev helper: C2 (uf250-1065 seed 0, 32 768 cubes) kernel time for a few option sets given as 'k=v,k=v' arguments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat
offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)
for arg in (sys.argv[1:] or [""]):
    opts = {k: int(v) for k, v in (kv.split("=") for kv in arg.split(",") if kv)}
    with g.Solver(250, pre.offsets, pre.lits, **opts) as s:
        s.set_cubes(cubes)
        ms, imp = [], 0
        for r in range(int(os.environ.get("REPS", "7"))):
            v, m, st = s.solve()
            if r >= 2:
                ms.append(st["kernel_ms"]); imp = st["implications"]
        print(f"{arg or 'default':32s} verdict {v} warps/block {st['warps_per_block']} regs? smem {st.get('smem_bytes')} kernel_ms median {np.median(ms):.2f} min {min(ms):.2f} "
              f"implications {imp} busy {st['warp_busy_frac']:.3f} splits {st['splits']}", flush=True)
