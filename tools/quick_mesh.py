#!/usr/bin/env python
"""Dev helper (one GPU): (a) the shard one rank of an N-GPU strong-scaling run sees — every N-th cube of C2 on the whole
GPU — to tune how fast dynamic splitting fills 2960 warps from 4096/N cubes; (b) a mesh of R handles sharing this GPU
(gpsat_multi_* with devices=[0]*R, 148/R blocks each): the cross-rank code path without a second GPU.
usage: python tools/quick_mesh.py "stride=8 split_gap=8" "mesh=2" "mesh=4 share_learnts=1 share_max_len=2" ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpupsat_b200 as g
from gpupsat_b200.instances import random_ksat

offs, lits = random_ksat(250, 1065, 0)
pre = g.Cnf.from_arrays(offs, lits).preprocess()
cubes = pre.choose_cubes(8, 32)
for spec in sys.argv[1:] or [""]:
    kvs = dict(kv.split("=") for kv in spec.split())
    stride, mesh, gpus = int(kvs.pop("stride", 1)), int(kvs.pop("mesh", 0)), int(kvs.pop("gpus", 0))
    if gpus:          # real GPUs of this box in one process (gpsat_multi_*: one host thread per GPU)
        mesh = gpus
    opts = {k: int(v) for k, v in kvs.items()}
    try:
        if mesh:
            mk = dict(n_gpus=gpus) if gpus else dict(n_gpus=mesh, devices=[0] * mesh, blocks=148 // mesh)
            with g.MultiSolver(250, pre.offsets, pre.lits, **mk, **opts) as s:
                s.set_cubes(cubes[::stride])
                ms, wall = [], []
                for r in range(4):
                    t = time.perf_counter()
                    v, m, st = s.solve()
                    if r:
                        ms.append(st["kernel_ms"])
                        wall.append(1e3 * (time.perf_counter() - t))
                ok = (s.job_records()["status"] == g.UNSAT).all()
            print(f"[{spec}] mesh x{mesh}: kernel ms {min(ms):.2f} / {sum(ms) / len(ms):.2f} wall {min(wall):.2f} impl {st['implications']:.3e} "
                  f"busy {st['warp_busy_frac']:.2f} splits {st['splits']} steals {st['steals']} foreign {st['foreign_clauses']} "
                  f"reduce {st['reduce_backend']} verdict {v} closed {ok}", flush=True)
        else:
            with g.Solver(250, pre.offsets, pre.lits, **opts) as s:
                s.set_cubes(cubes[::stride])
                ms = []
                for r in range(4):
                    v, m, st = s.solve()
                    if r:
                        ms.append(st["kernel_ms"])
                hist = s.debug_words()[72:88].tolist()
            print(f"[{spec}] children by log2(conflicts+1): {hist}")
            print(f"[{spec}] {len(cubes[::stride])} cubes: ms {min(ms):.2f} / {sum(ms) / len(ms):.2f} impl {st['implications']:.3e} "
                  f"confl {st['conflicts']} busy {st['warp_busy_frac']:.2f} splits {st['splits']} "
                  f"warps {st['blocks']}x{st['warps_per_block']} verdict {v}", flush=True)
    except Exception as e:
        print(f"[{spec}] FAILED {e}", flush=True)
