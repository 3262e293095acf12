#!/usr/bin/env python
"""Dev helper: samples the queue counters from a second host thread while one C2 solve runs (idle warps over time)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1          # every stride-th cube of C2 (the shard of one of `stride` GPUs)
opts = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[2:])}
offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)[0::stride]
s = g.Solver(250, pre.offsets, pre.lits, **opts)
s.set_cubes(cubes)
s.solve()                                   # warm-up
rows, stop = [], threading.Event()
def poll():
    t0 = time.perf_counter()
    while not stop.is_set():
        c = s.debug_ctrl()
        rows.append((1e3 * (time.perf_counter() - t0), int(c[0]), int(c[4]), int(c[5]), int(c[6]), int(c[7])))
        time.sleep(0.0005)
s.solve_begin()
th = threading.Thread(target=poll); th.start()
time.sleep(0.002)
t = time.perf_counter()
s.solve_step(0.0)
dt = 1e3 * (time.perf_counter() - t)
stop.set(); th.join()
v, m, st = s.solve_end()
rec = s.job_records()
c = np.sort(rec["conflicts"])[::-1]
print("per-root conflicts (whole subtree): max", c[:8].tolist(), "median", int(np.median(c)), "mean", round(float(c.mean()), 1),
      "share of the 1% hardest roots", round(float(c[: max(1, len(c) // 100)].sum() / c.sum()), 3), "splits of the hardest", np.sort(rec["reserved"])[::-1][:8].tolist())
print("step ms", round(dt, 2), "kernel", round(st["kernel_ms"], 2), "busy", round(st["warp_busy_frac"], 3), "splits", st["splits"], "warps", st["blocks"] * st["warps_per_block"])
print("   ms   next_job  pushed  popped  outstanding  idle")
last = -1
for r in rows:
    if r[0] - last >= 1.0:
        print(f"{r[0]:6.1f} {r[1]:9d} {r[2]:7d} {r[3]:7d} {r[4]:12d} {r[5]:5d}")
        last = r[0]
