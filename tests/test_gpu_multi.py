"""-m gpu, needs >= 2 GPUs: the real multi-GPU path (torchrun, NCCL all-gather of the exchange blocks over NVLink)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, n, m, seed, share_len, port, mode="nccl"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           str(n), str(m), str(seed), str(share_len), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("share_len", [0, 8])
def test_two_gpus_unsat_every_cube_closed(share_len):
    r = _run(2, 250, 1065, 0, share_len, 29541 + share_len)
    assert r["verdict"] == 1 and r["cubes"] == 8192 and r["cubes_closed_unsat"] == 8192
    if share_len:
        assert r["foreign_clauses_all_ranks"] > 0          # learnt clauses crossed the NVLink all-gather


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_gpus_sat_early_termination():
    r = _run(2, 200, 820, 1, 2, 29551)
    assert r["verdict"] == 0 and r["model_ok"] is True


# ---- the mesh over real peers (CUDA IPC between the torchrun processes, NVLink peer memory) -------------------------
@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("share_len", [0, 8])
def test_mesh_two_gpus_unsat_every_cube_closed(share_len):
    r = _run(2, 250, 1065, 0, share_len, 29561 + share_len, mode="mesh")
    assert r["verdict"] == 1 and r["cubes"] == 8192 and r["cubes_closed_unsat"] == 8192
    if share_len:
        assert r["foreign_clauses_all_ranks"] > 0          # learnt clauses were stored into the peer's pool over NVLink


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_mesh_two_gpus_sat_early_termination():
    r = _run(2, 200, 820, 1, 2, 29571, mode="mesh")
    assert r["verdict"] == 0 and r["model_ok"] is True


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_multi_solver_in_process_over_two_gpus():
    import gpupsat_b200 as g
    from gpupsat_b200.instances import random_ksat
    offs, lits = random_ksat(250, 1065, 0)
    pre = g.Cnf.from_arrays(offs, lits).preprocess()
    cubes = pre.choose_cubes(8, 32)
    with g.MultiSolver(pre.n_vars, pre.offsets, pre.lits, n_gpus=2) as ms:
        ms.set_cubes(cubes)
        verdict, model, stats = ms.solve()
        rec = ms.job_records()
    assert verdict == g.UNSAT and (rec["status"] == g.UNSAT).all() and stats["jobs_done"] == len(cubes)
    assert stats["reduce_backend"] in ("nccl", "host")
