"""N > 1 host logic on CPU: world_size-2 `gloo` runs of gpupsat_b200.multi_gpu.solve_sharded with a scripted stand-in
for the device solver (same five methods, CPU tensors, the exchange-block format of include/gpsat.h).  What is under
test is the protocol: static cube sharding, one all-gather per epoch, clause routing to the other rank only, early
termination when another rank holds a model, model broadcast, global verdict."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpupsat_b200 import multi_gpu as mg


class ScriptedSolver:
    """Finishes at epoch `finish_at` with `final`; publishes `clauses_per_epoch` clauses [rank, epoch, i] per epoch."""

    def __init__(self, rank, n_vars, finish_at, final, clauses_per_epoch=2):
        self.rank, self.n_vars, self.finish_at, self.final = rank, n_vars, finish_at, final
        self.cpe = clauses_per_epoch
        self.epoch, self.stopped, self.foreign, self.began = 0, False, [], False

    def solve_begin(self):
        self.began = True

    def solve_step(self, budget_ms):
        assert self.began and budget_ms > 0
        self.epoch += 1
        done = self.epoch >= self.finish_at
        return done, (self.final if done else mg.UNDEF)

    def exchange_pack(self, block, rank, done, verdict):
        b = block.numpy()
        b[:] = 0
        n = 0 if self.epoch > self.finish_at else self.cpe
        b[:8] = [mg.MAGIC, verdict, int(done), n * mg.SLOT_WORDS, n, rank, 0, 0]
        for i in range(n):
            b[8 + 16 * i: 8 + 16 * i + 4] = [3, 2 * self.rank, 2 * self.epoch + 1, 2 * i]

    def exchange_unpack(self, gathered, n_ranks, rank):
        g = gathered.numpy().reshape(n_ranks, -1)
        assert (g[:, 0] == mg.MAGIC).all() and (g[:, 5] == np.arange(n_ranks)).all()
        sat = [r for r in range(n_ranks) if g[r, 1] == mg.SAT]
        imported = 0
        for r in range(n_ranks):
            if r == rank:
                continue
            for i in range(int(g[r, 4])):
                self.foreign.append(tuple(g[r, 8 + 16 * i + 1: 8 + 16 * i + 4].tolist()))
                imported += 1
        fin = g[:, 2] != 0
        return {"sat_rank": sat[0] if sat else -1, "all_done": bool(fin.all()),
                "any_undef": bool(((g[:, 1] == mg.UNDEF) & fin).any()), "imported_clauses": imported,
                "jobs_done": int(g[:, 6].sum())}

    def request_stop(self):
        self.stopped = True

    def solve_end(self):
        model = np.full(self.n_vars, self.rank, dtype=np.uint8)
        v = self.final if self.epoch >= self.finish_at else mg.UNDEF
        return v, model, {"kernel_ms": 0.0}


def _worker(rank, world, port, scenario, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        finish_at, final = scenario[rank]
        s = ScriptedSolver(rank, 6, finish_at, final)
        verdict, model, stats, info = mg.solve_sharded(s, dist, rank, world, "cpu", budget_ms=5.0,
                                                       max_clauses_per_epoch=4)
        out[rank] = {"verdict": verdict, "model": None if model is None else model.tolist(), "epochs": info["epochs"],
                     "imported": info["imported_clauses"], "foreign": s.foreign, "stopped": s.stopped,
                     "bytes": info["exchange_bytes_per_epoch"]}
    finally:
        dist.destroy_process_group()


def _run(scenario):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, port, scenario, out), nprocs=2, join=True)
        return dict(out)


def test_unsat_needs_every_rank_and_routes_clauses():
    out = _run({0: (2, mg.UNSAT), 1: (4, mg.UNSAT)})
    assert out[0]["verdict"] == out[1]["verdict"] == mg.UNSAT
    assert out[0]["epochs"] == out[1]["epochs"] == 4             # the slower rank decides when the run ends
    assert out[0]["model"] is None
    # rank 0 published during epochs 1-2, rank 1 during 1-4; each rank received exactly the other's clauses
    assert out[1]["foreign"] == [(0, 2 * e + 1, 2 * i) for e in (1, 2) for i in (0, 1)]
    assert out[0]["foreign"] == [(2, 2 * e + 1, 2 * i) for e in (1, 2, 3, 4) for i in (0, 1)]
    assert out[0]["bytes"] == 4 * 2 * mg.block_words(4)


def test_sat_on_one_rank_stops_the_other_and_broadcasts_the_model():
    out = _run({0: (50, mg.UNSAT), 1: (3, mg.SAT)})
    assert out[0]["verdict"] == out[1]["verdict"] == mg.SAT
    assert out[0]["epochs"] == out[1]["epochs"] == 3
    assert out[0]["stopped"] and not out[1]["stopped"]          # early-termination flag reached rank 0
    assert out[0]["model"] == out[1]["model"] == [1] * 6          # rank 1's model on both ranks


def test_undef_on_a_finished_rank_makes_the_run_undef():
    out = _run({0: (1, mg.UNDEF), 1: (2, mg.UNSAT)})
    assert out[0]["verdict"] == out[1]["verdict"] == mg.UNDEF


def test_shard_cubes_is_a_partition():
    cubes = np.arange(4096 * 12).reshape(4096, 12)
    for world in (1, 2, 4, 8):
        parts = [mg.shard_cubes(cubes, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == 4096
        assert np.array_equal(np.sort(np.concatenate(parts)[:, 0]), cubes[:, 0])
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
