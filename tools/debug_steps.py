#!/usr/bin/env python
"""Dev helper: budgeted steps with a watchdog thread that dumps the queue counters if a step does not return."""
import sys, os, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
n, m, seed = 150, 639, 2
offs, lits = random_ksat(n, m, seed)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(1, 4)[0::2]
s = g.Solver(n, pre.offsets, pre.lits, stop_on_sat=0, share_learnts=1, share_max_len=8)
s.set_cubes(cubes)
s.solve_begin()
state = {"in_step": False, "t": 0.0, "epoch": 0}

def dog():
    while True:
        time.sleep(1.0)
        if state["in_step"] and time.time() - state["t"] > 3.0:
            for i in range(3):
                print("WATCHDOG epoch", state["epoch"], "ctrl", s.debug_ctrl().tolist(), flush=True)
                time.sleep(0.5)
            print("requesting stop", flush=True)
            s.request_stop()
            time.sleep(3.0)
            print("after stop ctrl", s.debug_ctrl().tolist(), flush=True)
            os._exit(3)

threading.Thread(target=dog, daemon=True).start()
for epoch in range(3000):
    state.update(in_step=True, t=time.time(), epoch=epoch)
    done, verdict = s.solve_step(budget_ms=budget)
    state["in_step"] = False
    if epoch < 5 or epoch % 50 == 0:
        print("epoch", epoch, "done", done, "verdict", verdict, "ctrl", s.debug_ctrl().tolist(), "kernel_ms", round(s.last_kernel_ms(), 3), flush=True)
    if done:
        break
v, model, st = s.solve_end()
print("finished: epochs", epoch + 1, "verdict", v, {k: st[k] for k in ("jobs_done", "conflicts", "splits", "kernel_ms", "kernel_launches")})
