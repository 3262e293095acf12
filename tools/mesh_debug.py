#!/usr/bin/env python
"""Dev helper (one GPU): R mesh ranks as separate handles on this GPU, per-rank statistics and control-block words.
usage: python tools/mesh_debug.py R "opt=val ..." ["opt=val ..." ...]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat

R = int(sys.argv[1])
offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)
for spec in sys.argv[2:] or [""]:
    opts = {k: int(v) for k, v in (kv.split("=") for kv in spec.split())}
    solvers = [g.Solver(250, pre.offsets, pre.lits, device=0, blocks=148 // R, **opts) for _ in range(R)]
    for s in solvers:
        s.set_cubes(cubes)
    if R > 1:
        g.mesh_attach_local(solvers)
    out = [None] * R
    for rep in range(3):
        bar = threading.Barrier(R)

        def run(r):
            s = solvers[r]
            s.solve_begin()
            bar.wait()
            t = time.perf_counter()
            done = False
            while not done:
                done, _ = s.solve_step(3000.0)
            out[r] = (1e3 * (time.perf_counter() - t),) + s.solve_end() + (s.debug_words(),)

        rows, stop = [], threading.Event()

        def poll():
            t0 = time.perf_counter()
            while not stop.is_set():
                cs = [s.debug_ctrl() for s in solvers]
                rows.append((1e3 * (time.perf_counter() - t0), int(cs[0][0])) + tuple(int(c[7]) for c in cs) + tuple(int(c[6]) for c in cs)
                            + tuple(int(c[4] - c[5]) for c in cs) + tuple(int(c[10]) for c in cs))
                time.sleep(0.0005)

        th = [threading.Thread(target=run, args=(r,)) for r in range(R)]
        pt = threading.Thread(target=poll)
        if rep == 2:
            pt.start()
        [t.start() for t in th]
        [t.join() for t in th]
        if rep == 2:
            stop.set()
            pt.join()
    print(f"[{spec}] ranks {R}")
    print("     ms  cursor  idle/rank  open/rank  queued/rank  steals/rank")
    last = -9
    for row in rows:
        if os.environ.get('TL') and row[0] - last >= 3.0:
            print("  ", " ".join(f"{x:7.1f}" if i == 0 else f"{x:6d}" for i, x in enumerate(row)))
            last = row[0]
    for r in range(R):
        wall, v, m, st, w = out[r]
        print(f"   rank {r}: wall {wall:.1f} kernel {st['kernel_ms']:.1f} launches {st['kernel_launches']} busy {st['warp_busy_frac']:.2f} "
              f"impl(local) {st['implications']:.3e} confl {st['conflicts']} splits {st['splits']} steals {w[67]} remote tries {w[70]} "
              f"created {w[32]} closed {w[33]} cursor {w[24]} children hist {w[72:82].tolist()}")
        ph = w[104:116].view(np.int64)
        ph2 = w[116:128].view(np.int64)
        print("       STOLEN jobs: conflicts", int(w[71]), "phase warp-ms", [round(float(x) * 1e-6, 1) for x in ph2])
        print("       phase warp-ms [reset, import, propagate, analyze, split, reduce_db]:", [round(float(x) * 1e-6, 1) for x in ph],
              " job warp-ms", round(st['warp_busy_frac'] * st['kernel_ms'] * st['blocks'] * st['warps_per_block'], 1))
    for s in solvers:
        s.close()
