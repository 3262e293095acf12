"""Cube-and-conquer over the GPUs of one box: one process per GPU, cubes sharded statically, one collective per epoch.

There is no reference equivalent (the reference is single-GPU); what is kept from it is the unit of
work — a JobChooser cube (JobsManager/JobChooser.cu:75-90) pulled from an atomic cursor (SATSolver/JobsQueue.cu:10-32)
— and the meaning of the early-termination flag (`*state`, SATSolver/main.cu:259-269), which here travels inside the
header of the exchange block instead of being polled between kernel launches.

Per epoch and per rank:
    solver.solve_step(budget_ms)            persistent kernel until the budget is spent (unfinished cubes park themselves)
    solver.exchange_pack(block, ...)        header [magic, verdict, done, ...] + short learnt clauses published since the
                                            last pack, written by a CUDA kernel straight into `block` (device memory)
    all_gather(block) -> gathered           ONE collective: NCCL over NVLink on the GPU box, gloo in the CPU tests
    solver.exchange_unpack(gathered, ...)   foreign clauses -> this GPU's foreign pool; flags reduced from the headers
The data path has no other collective: the clause database is replicated, cubes never move.

`solver` is a gpupsat_b200.Solver (or, in the CPU tests, an object with the same five methods working on CPU
tensors); `dist` is torch.distributed or None for a single process.

`solve_mesh` is the fused form (the product path on a GPU box): the GPUs are joined in a mesh over NVLink peer memory
(include/gpsat.h: gpsat_mesh_*), a solve is ONE persistent launch per GPU inside which idle warps take split-off cubes
from the other GPUs' rings, learnt clauses are stored straight into the peers' pools and termination is detected —
and the only collectives left are the all-gather of the 64-byte IPC handles when the mesh is formed and the
all-reduces of the per-cube result block at the end."""
from __future__ import annotations

import numpy as np

SAT, UNSAT, UNDEF = 0, 1, 2
HEADER_WORDS = 8           # GPSAT_XCHG_HEADER_WORDS
SLOT_WORDS = 16            # GPSAT_POOL_SLOT_WORDS
MAGIC = 0x47505358         # GPSAT_XCHG_MAGIC


def shard_cubes(cubes, rank: int, world: int):
    """cube j -> rank j mod world (interleaved: neighbouring cubes differ in their last literals and tend to be equally
    hard, so every rank gets the same mix)."""
    return cubes[rank::world]


def block_words(max_clauses_per_epoch: int) -> int:
    return HEADER_WORDS + max_clauses_per_epoch * SLOT_WORDS


def _all_gather(dist, gathered, block, world):
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(gathered, block)
    else:
        dist.all_gather(list(gathered.view(world, -1).unbind(0)), block)


def solve_sharded(solver, dist, rank: int, world: int, device, *, budget_ms: float = 20.0,
                  max_clauses_per_epoch: int = 1024, max_epochs: int | None = None):
    """Runs the epoch loop until some rank finds a model, every rank has closed all its cubes, or max_epochs is
    reached.  Returns (verdict, model or None, stats, info): the verdict is global (identical on every rank), the model
    is broadcast from the lowest rank that holds one, stats are this rank's, info counts epochs / exchanged clauses."""
    import torch

    words = block_words(max_clauses_per_epoch)
    block = torch.zeros(words, dtype=torch.int32, device=device)
    gathered = torch.zeros(world * words, dtype=torch.int32, device=device)
    is_cuda = torch.device(device).type == "cuda"
    solver.solve_begin()
    epochs, imported, sat_rank, all_done, any_undef, jobs = 0, 0, -1, False, False, 0
    while True:
        done, verdict = solver.solve_step(budget_ms)
        solver.exchange_pack(block, rank, done, verdict)
        if dist is not None and world > 1:
            _all_gather(dist, gathered, block, world)
            if is_cuda:
                torch.cuda.current_stream(device).synchronize()   # unpack runs on the library's own stream
        else:
            gathered.copy_(block)
        r = solver.exchange_unpack(gathered, world, rank)
        epochs += 1
        imported += r["imported_clauses"]
        sat_rank, all_done, any_undef, jobs = r["sat_rank"], r["all_done"], r["any_undef"], r["jobs_done"]
        if sat_rank >= 0:
            if verdict != SAT:
                solver.request_stop()          # early termination: another GPU holds a model
            break
        if all_done or (max_epochs is not None and epochs >= max_epochs):
            break
    local_verdict, model, stats = solver.solve_end()
    if sat_rank >= 0:
        verdict = SAT
        m = torch.as_tensor(np.ascontiguousarray(model, dtype=np.uint8)).to(device)
        if dist is not None and world > 1:
            dist.broadcast(m, src=sat_rank)
        model = m.cpu().numpy()
    else:
        verdict = UNSAT if (all_done and not any_undef) else UNDEF
        model = None
    info = {"epochs": epochs, "imported_clauses": imported, "exchange_bytes_per_epoch": 4 * words * world,
            "jobs_done_all_ranks": jobs, "local_verdict": local_verdict}
    return verdict, model, stats, info


# ---------------------------------------------------------------------------------------------------------------
# mesh: the GPUs of one box as ONE work pool (one process per GPU, CUDA IPC)
# ---------------------------------------------------------------------------------------------------------------
def mesh_join(solver, dist, rank: int, world: int, device):
    """Forms the mesh: every rank exports its queue region (a 64-byte CUDA IPC handle), ONE all-gather distributes
    the handles, every rank maps the others' regions.  Every rank's solver holds the SAME, complete cube list: the
    cubes are handed out by one cursor (rank 0's) that all GPUs advance over NVLink."""
    import torch

    mine = torch.as_tensor(solver.mesh_export()).to(device)
    allh = torch.zeros(world * 64, dtype=torch.uint8, device=device)
    if dist is not None and world > 1:
        _all_gather(dist, allh, mine, world)
    else:
        allh.copy_(mine)
    solver.mesh_attach_ipc(world, rank, allh.cpu().numpy())
    return mesh_result_block(solver, device)


def mesh_result_block(solver, device):
    import torch

    return torch.zeros(solver.mesh_result_words(), dtype=torch.int32, device=device)


def reduce_results(dist, block, n_roots: int, world: int):
    """[flags n_roots | open descendants n_roots | records n_roots x 20 words]: MAX, SUM (int32), SUM (int64)."""
    if dist is None or world <= 1:
        return
    dist.all_reduce(block[:n_roots], op=dist.ReduceOp.MAX)
    dist.all_reduce(block[n_roots:2 * n_roots], op=dist.ReduceOp.SUM)
    import torch

    dist.all_reduce(block[2 * n_roots:].view(torch.int64), op=dist.ReduceOp.SUM)


def solve_mesh(solver, dist, rank: int, world: int, device, block, n_roots: int, *, budget_ms: float = 0.0,
               max_steps: int | None = None):
    """One mesh solve.  Returns (verdict, model or None, global stats, info).  budget_ms = 0: one launch per GPU until
    the whole job is done; > 0: time-bounded steps (unfinished cubes park in place; used for the time-bounded configs)."""
    import time
    import torch

    tm = [time.perf_counter()]
    solver.solve_begin()
    tm.append(time.perf_counter())
    if dist is not None and world > 1:
        dist.barrier()                    # nobody steals from a ring its owner has not reset yet
    tm.append(time.perf_counter())
    steps = 0
    while True:
        done, _ = solver.solve_step(budget_ms)
        steps += 1
        if done or (max_steps is not None and steps >= max_steps):
            break
    tm.append(time.perf_counter())
    local_verdict, model, local_stats = solver.solve_end()
    solver.mesh_results_pack(block)
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    tm.append(time.perf_counter())
    reduce_results(dist, block, n_roots, world)
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()       # unpack runs on the library's own stream
    tm.append(time.perf_counter())
    verdict, stats = solver.mesh_results_unpack(block)
    tm.append(time.perf_counter())
    for k in ("kernel_ms", "kernel_launches", "warp_busy_frac", "steals", "foreign_clauses", "pool_clauses", "blocks",
              "warps_per_block", "smem_bytes_per_block", "state_in_smem"):
        stats[k] = local_stats[k]
    sat_rank = -1
    if verdict == SAT:
        has = torch.tensor([rank if local_verdict == SAT else world], dtype=torch.int32, device=device)
        if dist is not None and world > 1:
            dist.all_reduce(has, op=dist.ReduceOp.MIN)
        sat_rank = int(has[0])
        m = torch.as_tensor(np.ascontiguousarray(model, dtype=np.uint8)).to(device)
        if dist is not None and world > 1:
            dist.broadcast(m, src=sat_rank)
        model = m.cpu().numpy()
    else:
        model = None
    phases = dict(zip(("begin", "barrier", "steps", "end_pack", "reduce", "unpack"),
                      (1e3 * (b - a) for a, b in zip(tm, tm[1:]))))     # host milliseconds of this rank
    return verdict, model, stats, {"steps": steps, "sat_rank": sat_rank, "local_verdict": local_verdict, "host_ms": phases}
