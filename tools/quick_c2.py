#!/usr/bin/env python
"""Dev helper: C2 solve (uf250-1065 seed 0, 4096 cubes) under several option sets, kernel time only (no CPU legs).
usage: python tools/quick_c2.py "split_gap=8 split_burst=4" "warps_per_block=16" ...   (an empty spec = defaults)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat

offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)
for spec in sys.argv[1:] or [""]:
    kvs = [kv.split("=") for kv in spec.split()]
    for k in [k for k in os.environ if k.startswith("GPSAT_") and k != "GPSAT_BENCH_C4"]:
        del os.environ[k]
    os.environ.update({k: v for k, v in kvs if k.startswith("GPSAT_")})       # library knobs read through getenv
    opts = {k: int(v) for k, v in kvs if not k.startswith("GPSAT_")}
    try:
        with g.Solver(250, pre.offsets, pre.lits, **opts) as s:
            s.set_cubes(cubes)
            ms, st = [], None
            for r in range(4):
                v, m, st = s.solve()
                if r:
                    ms.append(st["kernel_ms"])
        print(f"[{spec}] ms {min(ms):.2f} / {sum(ms) / len(ms):.2f}  impl {st['implications']:.3e} confl {st['conflicts']} "
              f"busy {st['warp_busy_frac']:.2f} splits {st['splits']} warps {st['blocks']}x{st['warps_per_block']} verdict {v}", flush=True)
    except Exception as e:
        print(f"[{spec}] FAILED {e}", flush=True)
