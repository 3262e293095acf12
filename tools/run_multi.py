#!/usr/bin/env python
"""One instance over N GPUs (torchrun, one process per GPU): time-to-solve with the epoch exchange.
usage: torchrun --nproc-per-node N tools/run_multi.py N_VARS N_CLAUSES SEED SHARE_MAX_LEN [EPOCH_MS] [TIME_LIMIT_S]"""
import json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpupsat_b200 as g
from gpupsat_b200 import multi_gpu as mg
from gpupsat_b200.instances import check_model, random_ksat

n, m, seed, share_len = (int(x) for x in sys.argv[1:5])
epoch_ms = float(sys.argv[5]) if len(sys.argv) > 5 else 50.0
limit_s = float(sys.argv[6]) if len(sys.argv) > 6 else 60.0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
offs, lits = random_ksat(n, m, seed)
cnf = g.Cnf.from_arrays(offs, lits)
pre = cnf.preprocess()
cubes = pre.choose_cubes(8 * world, 32)
opts = dict(share_learnts=1, share_max_len=share_len) if share_len else {}
with g.Solver(cnf.n_vars, pre.offsets, pre.lits, device=local, **opts) as s:
    s.set_cubes(mg.shard_cubes(cubes, rank, world))
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        verdict, model, st, info = mg.solve_sharded(s, dist if world > 1 else None, rank, world, dev, budget_ms=epoch_ms,
                                                    max_clauses_per_epoch=4096, max_epochs=int(limit_s * 1e3 / epoch_ms))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    t = torch.tensor([wall, st["kernel_ms"], float(st["implications"]), float(st["conflicts"]), float(st["foreign_clauses"])],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    else:
        tmax = t
if rank == 0:
    print(json.dumps({"instance": f"uf{n}-{m} seed {seed}", "gpus": world, "cubes": len(cubes), "share_max_len": share_len,
                      "verdict": {0: "SAT", 1: "UNSAT", 2: "UNDEF"}[verdict], "wall_ms": 1e3 * float(tmax[0]),
                      "kernel_ms_max": float(tmax[1]), "implications": float(t[2]), "conflicts": float(t[3]),
                      "foreign_clauses": float(t[4]), "epochs": info["epochs"],
                      "implications_per_s": float(t[2]) / float(tmax[0])}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
