"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun (NCCL).  Solves one instance with the epoch
loop of gpupsat_b200.multi_gpu and prints a JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpupsat_b200 as g  # noqa: E402
from gpupsat_b200 import multi_gpu as mg  # noqa: E402
from gpupsat_b200.instances import check_model, random_ksat  # noqa: E402


def main():
    n, m, seed, share_len = (int(x) for x in sys.argv[1:5])
    mode = sys.argv[5] if len(sys.argv) > 5 else "nccl"              # nccl: epoch all-gather; mesh: NVLink peer memory
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    offs, lits = random_ksat(n, m, seed)
    cnf = g.Cnf.from_arrays(offs, lits)
    pre = cnf.preprocess()
    cubes = pre.choose_cubes(8 * world, 32)
    mine = mg.shard_cubes(cubes, rank, world)
    opts = dict(share_learnts=1, share_max_len=share_len) if share_len else {}
    steals = 0
    with g.Solver(cnf.n_vars, pre.offsets, pre.lits, device=local, **opts) as s:
        s.set_cubes(cubes if mode == "mesh" else mine)      # mesh: every rank holds all cubes (one root cursor)
        if mode == "mesh":
            block = mg.mesh_join(s, dist, rank, world, dev)
            verdict, model, stats, info = mg.solve_mesh(s, dist, rank, world, dev, block, len(cubes), budget_ms=2000.0,
                                                        max_steps=30)
            rec = s.job_records(len(cubes))                           # global records after the reduction
            info["epochs"] = info["steps"]
            steals = stats["steals"]
        else:
            verdict, model, stats, info = mg.solve_sharded(s, dist, rank, world, dev, budget_ms=5.0,
                                                           max_clauses_per_epoch=512)
            rec = s.job_records()
    ok = bool(check_model(pre.offsets, pre.lits, model)) if verdict == g.SAT else None
    counts = mode != "mesh" or rank == 0          # mesh: every rank holds the same (global) records — count them once
    closed = torch.tensor([int((rec["status"] == g.UNSAT).sum()) if counts else 0, len(rec) if counts else 0,
                           stats["foreign_clauses"], steals], device=dev)
    dist.all_reduce(closed)
    if rank == 0:
        print(json.dumps({"verdict": int(verdict), "model_ok": ok, "epochs": info["epochs"],
                          "cubes_closed_unsat": int(closed[0]), "cubes": int(closed[1]),
                          "foreign_clauses_all_ranks": int(closed[2]), "steals_all_ranks": int(closed[3]),
                          "world": world, "mode": mode}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
