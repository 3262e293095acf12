"""-m gpu: the mesh (several GPUs as ONE work pool over peer memory, include/gpsat.h gpsat_mesh_* / gpsat_multi_*).

Every test here runs on ONE GPU: the ranks of the mesh are separate handles whose persistent kernels share the GPU
(74 blocks each), and "peer memory" is then ordinary device memory — the code path (remote ring pops, staging of the
hand-off block, communication warp, termination detection, clause push, result reduction) is the one that runs across
NVLink.  tests/test_gpu_multi.py repeats them over real peers when the box has >= 2 GPUs."""
import threading

import numpy as np
import pytest

import gpupsat_b200 as g
from gpupsat_b200.instances import check_model, pigeonhole, random_ksat

pytestmark = pytest.mark.gpu


def _instance(n, m, seed, b=8, t=32):
    offs, lits = random_ksat(n, m, seed)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    assert pre.status == g.UNDEF
    return pre, pre.choose_cubes(b, t)


def _mesh_solve(pre, cubes, n_ranks=2, budget_ms=3000.0, rank_blocks=None, **opts):
    """every rank holds all cubes; rank_blocks = CTAs per rank (default: the GPU split evenly)"""
    import torch
    n_roots = len(cubes)
    rank_blocks = rank_blocks or [148 // n_ranks] * n_ranks
    solvers = [g.Solver(pre.n_vars, pre.offsets, pre.lits, device=0, blocks=rank_blocks[r], **opts) for r in range(n_ranks)]
    try:
        for s in solvers:
            s.set_cubes(cubes)
        g.mesh_attach_local(solvers)
        blocks = [torch.zeros(s.mesh_result_words(), dtype=torch.int32, device="cuda:0") for s in solvers]
        out = [None] * n_ranks
        barrier = threading.Barrier(n_ranks)

        def run(r):
            s = solvers[r]
            s.solve_begin()
            barrier.wait()
            for _ in range(20):
                done, _v = s.solve_step(budget_ms)
                if done:
                    break
            out[r] = (done,) + s.solve_end()
            s.mesh_results_pack(blocks[r])

        th = [threading.Thread(target=run, args=(r,)) for r in range(n_ranks)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert all(o is not None and o[0] for o in out), "a rank did not finish"
        nr = n_roots
        red = blocks[0].clone()
        for b in blocks[1:]:
            red[:nr] = torch.maximum(red[:nr], b[:nr])
            red[nr:2 * nr] += b[nr:2 * nr]
            red[2 * nr:].view(torch.int64).add_(b[2 * nr:].view(torch.int64))
        verdict, stats = solvers[0].mesh_results_unpack(red)
        rec = solvers[0].job_records(n_roots)
        return verdict, stats, rec, out
    finally:
        for s in solvers:
            s.close()


def test_mesh_small_rank_lives_off_the_big_one_and_everything_closes():
    pre, cubes = _instance(250, 1065, 0, 1, 4)            # 64 cubes for 2 x 1480 warps: almost every job is a split-off cube
    verdict, stats, rec, out = _mesh_solve(pre, cubes, rank_blocks=[100, 48])
    assert verdict == g.UNSAT
    assert len(rec) == len(cubes) and (rec["status"] == g.UNSAT).all()
    assert sum(o[3]["steals"] for o in out) > 0           # split-off cubes crossed between the ranks' rings
    assert stats["splits"] > 0 and stats["jobs_done"] == len(cubes)


@pytest.mark.parametrize("share", [0, 8])
def test_mesh_two_ranks_unsat_all_cubes_closed(share):
    pre, cubes = _instance(250, 1065, 0, 2, 32)
    opts = dict(share_learnts=1, share_max_len=share) if share else {}
    verdict, stats, rec, out = _mesh_solve(pre, cubes, **opts)
    assert verdict == g.UNSAT
    assert len(rec) == len(cubes) and (rec["status"] == g.UNSAT).all()
    assert stats["jobs_done"] == len(cubes) and stats["conflicts"] > 0
    if share:
        assert sum(o[3]["foreign_clauses"] for o in out) > 0      # clauses crossed between the ranks inside the launch


def test_mesh_sat_early_termination_and_model():
    offs, lits = random_ksat(200, 820, 1)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cubes = pre.choose_cubes(2, 32)
    verdict, stats, rec, out = _mesh_solve(pre, cubes)
    assert verdict == g.SAT
    models = [o[2] for o in out if o[1] == g.SAT]
    assert models and all(check_model(pre.offsets, pre.lits, m) for m in models)


def test_multi_solver_two_handles_on_one_gpu():
    pre, cubes = _instance(250, 1065, 0, 2, 32)
    with g.MultiSolver(pre.n_vars, pre.offsets, pre.lits, n_gpus=2, devices=[0, 0], blocks=74) as ms:
        ms.set_cubes(cubes)
        verdict, model, stats = ms.solve()
        rec = ms.job_records()
    assert verdict == g.UNSAT and (rec["status"] == g.UNSAT).all() and stats["jobs_done"] == len(cubes)
    # same formula, sequential cube (the reference's -b 1 -t 1 mode) through the same host
    offs, lits = pigeonhole(7, 6)
    pre2 = g.Cnf.from_arrays(offs, lits).preprocess()
    with g.MultiSolver(pre2.n_vars, pre2.offsets, pre2.lits, n_gpus=2, devices=[0, 0], blocks=74) as ms:
        ms.set_cubes(None)
        verdict, model, stats = ms.solve()
    assert verdict == g.UNSAT and stats["splits"] > 0          # one root cube: every other job is a split-off cube
