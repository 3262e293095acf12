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
tensors); `dist` is torch.distributed or None for a single process."""
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
