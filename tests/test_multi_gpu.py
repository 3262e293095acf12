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


# ---------------------------------------------------------------------------------------------------------------
# mesh harness (solve_mesh): join by one all-gather of the 64-byte handles, one launch, reduction of the result block
# [flags: MAX | open descendants: SUM int32 | records: SUM int64], global verdict, model from the lowest SAT rank
# ---------------------------------------------------------------------------------------------------------------
REC_WORDS = 20
FLAG_UNSAT, FLAG_SAT = 1, 5


class ScriptedMeshSolver:
    """Rank r closes the roots it owns (g mod world == r) plus, like a thief, root `steal` of the next rank; each
    closed root contributes flag UNSAT (or SAT for `sat_root`), -1/+1 open-descendant bookkeeping and 7 implications."""

    def __init__(self, rank, world, n_roots, n_vars, sat_root=None, leave_open=None):
        self.rank, self.world, self.n_roots, self.n_vars = rank, world, n_roots, n_vars
        self.sat_root, self.leave_open = sat_root, leave_open
        self.attached, self.began, self.steps = None, False, 0

    def mesh_export(self):
        return np.full(64, 10 + self.rank, dtype=np.uint8)

    def mesh_attach_ipc(self, n_ranks, rank, handles):
        h = np.asarray(handles).reshape(n_ranks, 64)
        assert (h[:, 0] == 10 + np.arange(n_ranks)).all()           # every rank's handle arrived, in rank order
        self.attached = (n_ranks, rank)

    def mesh_result_words(self):
        return self.n_roots * (2 + REC_WORDS)

    def solve_begin(self):
        assert self.attached == (self.world, self.rank)
        self.began = True

    def solve_step(self, budget_ms):
        assert self.began
        self.steps += 1
        return True, mg.UNDEF

    def solve_end(self):
        v = mg.SAT if (self.sat_root is not None and self.sat_root % self.world == self.rank) else mg.UNDEF
        return v, np.full(self.n_vars, self.rank, dtype=np.uint8), \
            {k: 0 for k in ("kernel_ms", "kernel_launches", "warp_busy_frac", "steals", "foreign_clauses", "pool_clauses",
                            "blocks", "warps_per_block", "smem_bytes_per_block", "state_in_smem")}

    def mesh_results_pack(self, block):
        b = block.numpy()
        b[:] = 0
        nr = self.n_roots
        rec = b[2 * nr:].view(np.int64).reshape(nr, REC_WORDS // 2)
        for g_ in range(nr):
            if g_ % self.world == self.rank:
                b[nr + g_] += 1                                    # the owner opened the root ...
                if g_ == self.leave_open:
                    continue
                b[nr + g_] += 1                                    # ... split a child off it, closed its own half
                b[nr + g_] -= 1
                b[g_] = FLAG_SAT if g_ == self.sat_root else FLAG_UNSAT
                rec[g_, 2] += 7
            if (g_ + 1) % self.world == self.rank and g_ != self.leave_open:
                b[nr + g_] -= 1                                    # the child was closed HERE, on another rank
                b[g_] = max(b[g_], FLAG_UNSAT)
                rec[g_, 2] += 7

    def mesh_results_unpack(self, block):
        b = block.numpy()
        nr = self.n_roots
        flags, pend = b[:nr], b[nr:2 * nr]
        rec = b[2 * nr:].view(np.int64).reshape(nr, REC_WORDS // 2)
        if (flags == FLAG_SAT).any():
            v = mg.SAT
        elif ((flags == FLAG_UNSAT) & (pend == 0)).all():
            v = mg.UNSAT
        else:
            v = mg.UNDEF
        return v, {"implications": int(rec[:, 2].sum()), "jobs_done": int(((flags > 0) & (pend == 0)).sum())}


def _mesh_worker(rank, world, port, kw, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_roots = 9
        s = ScriptedMeshSolver(rank, world, n_roots, 5, **kw)
        block = mg.mesh_join(s, dist, rank, world, "cpu")
        verdict, model, stats, info = mg.solve_mesh(s, dist, rank, world, "cpu", block, n_roots)
        out[rank] = {"verdict": verdict, "model": None if model is None else model.tolist(), "stats": stats,
                     "sat_rank": info["sat_rank"], "steps": info["steps"]}
    finally:
        dist.destroy_process_group()


def _run_mesh(**kw):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_mesh_worker, args=(2, port, kw, out), nprocs=2, join=True)
        return dict(out)


def test_mesh_unsat_needs_the_contributions_of_every_rank():
    out = _run_mesh()
    assert out[0]["verdict"] == out[1]["verdict"] == mg.UNSAT
    # every root was closed half by its owner and half by the other rank: only the reduced block says "closed"
    assert out[0]["stats"]["jobs_done"] == out[1]["stats"]["jobs_done"] == 9
    assert out[0]["stats"]["implications"] == 9 * 14
    assert out[0]["model"] is None and out[0]["steps"] == 1


def test_mesh_open_descendant_on_one_rank_keeps_the_verdict_undef():
    out = _run_mesh(leave_open=4)
    assert out[0]["verdict"] == out[1]["verdict"] == mg.UNDEF


def test_mesh_sat_model_comes_from_the_lowest_sat_rank():
    out = _run_mesh(sat_root=3)                                      # root 3 belongs to rank 1
    assert out[0]["verdict"] == out[1]["verdict"] == mg.SAT
    assert out[0]["sat_rank"] == out[1]["sat_rank"] == 1
    assert out[0]["model"] == out[1]["model"] == [1] * 5
