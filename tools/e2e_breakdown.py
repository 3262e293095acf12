#!/usr/bin/env python
"""Dev helper: where the end-to-end (host buffers -> verdict) time of one C2 solve goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat
offs, lits = random_ksat(250, 1065, 0)
cnf = g.Cnf.from_arrays(offs, lits)
pre = cnf.preprocess()
cubes = pre.choose_cubes(8, 32)
keep = g.Solver(cnf.n_vars, pre.offsets, pre.lits)      # a second live handle, as in bench.py
keep.set_cubes(cubes); keep.solve()
for it in range(6):
    t0 = time.perf_counter()
    s = g.Solver(cnf.n_vars, pre.offsets, pre.lits)
    t1 = time.perf_counter()
    s.set_cubes(cubes)
    t2 = time.perf_counter()
    s.solve_begin()
    t3 = time.perf_counter()
    done, v = s.solve_step(0.0)
    t4 = time.perf_counter()
    verdict, model, st = s.solve_end()
    t5 = time.perf_counter()
    s.close()
    t6 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.2f} set_cubes {1e3*(t2-t1):.2f} begin {1e3*(t3-t2):.2f} step {1e3*(t4-t3):.2f} "
          f"(kernel {st['kernel_ms']:.2f}) end {1e3*(t5-t4):.2f} close {1e3*(t6-t5):.2f} total {1e3*(t6-t0):.2f} ms")
