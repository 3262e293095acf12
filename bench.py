#!/usr/bin/env python
"""bench.py — the driver's measurement contract for gpupsat_b200.

Workload (BASELINE.json configs[1]): uniform random 3-SAT n=250 m=1065 (r=4.26, splitmix64 seed 0, UNSAT) split into
4096 JobChooser cubes (reference semantics, -b 8 -t 32 -> k=12).  A step = one complete cube-and-conquer solve of
that instance: every cube is refuted by warp-per-cube CDCL (BCP + 1-UIP learning + VSIDS + geometric restarts).
  value      BCP implications/s over the whole job, formula and cubes already resident in HBM, CUDA-event time of
             the solve kernel on the library's stream
  e2e        same metric through the C ABI from HOST buffers: gpsat_create (H2D of the formula index) +
             gpsat_set_cubes (H2D) + gpsat_solve (kernel, D2H of verdict/model/per-cube records) + gpsat_destroy
  roofline   algorithmic bytes of the solve kernel (8 B per watch entry visited + 4 B per clause word
             read + 8 B per implication, from exact in-kernel counters) / kernel time, against the measured HBM peak
  cpu_baseline  the reference's own solver classes built for the host (oracle/_ref) on a bounded sample of the same
             cubes, one core (N=1, rank 0 only)
N > 1 (torchrun, one process per GPU): weak scaling — 4096 cubes per GPU (k = 12 + log2 N over the same formula),
cube j -> rank j mod N, no data-path collective; per epoch one NCCL all-reduce of the early-termination flag.

`--impl reference` times the reference's CPU implementation (oracle/_ref, all host cores, one process per core over
slices of a bounded cube sample) on the same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VARS, N_CLAUSES, SEED = 250, 1065, 0
METRIC = "bcp_implications_per_sec"
UNIT = "implications/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_workload(n_gpus):
    import gpupsat_b200 as g
    from gpupsat_b200.instances import random_ksat
    offs, lits = random_ksat(N_VARS, N_CLAUSES, SEED)
    cnf = g.Cnf.from_arrays(offs, lits)
    pre = cnf.preprocess()
    assert pre.status == g.UNDEF
    strong = os.environ.get("GPSAT_BENCH_SCALING", "weak") == "strong"
    cubes = pre.choose_cubes(8 if strong else 8 * n_gpus, 32)   # 10*B*T jobs wanted -> k = 12 (+ log2(n_gpus) when weak)
    return cnf, pre, cubes


def algorithmic_bytes(stats):
    return 8 * stats["watchers_visited"] + 4 * stats["clause_words_read"] + 8 * stats["implications"]


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (reference's own classes on host cores).  The only place besides tests/ that executes oracle/.
# ---------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    offs, lits, cubes = args
    from oracle.binding import Reference
    from oracle.binding import Quiet
    with Quiet():
        R = Reference(offs, lits)
        R.reset_counters()
        t = time.perf_counter()
        undef = 0
        for c in cubes:
            st, _ = R.solve(c)
            undef += st == 2
        dt = time.perf_counter() - t
        imp, dec = R.counters()
    return imp, dec, dt, undef, len(cubes)


def cpu_reference_sample(pre_offs, pre_lits, cubes, n_sample, procs):
    """Reference solver (as shipped: MAX_ITERATIONS 1000 -> UNDEF on hard cubes) on the first n_sample cubes."""
    from oracle.binding import REF_DIR
    kind = "reference"
    if not os.path.exists(os.path.join(REF_DIR, "libgpsat_ref.so")):
        kind = "port"
    sample = cubes[:n_sample]
    t0 = time.perf_counter()
    if kind == "reference":
        if procs <= 1:
            res = [_ref_worker((pre_offs, pre_lits, sample))]
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_ref_worker, [(pre_offs, pre_lits, sample[i::procs]) for i in range(procs)])
        imp = sum(r[0] for r in res)
        undef = sum(r[3] for r in res)
    else:
        from oracle.binding import Oracle
        o = Oracle(N_VARS, pre_offs, pre_lits)
        k = sample.shape[1]
        r = o.run(np.arange(0, sample.size + 1, k, dtype=np.int64), sample.reshape(-1), stop_on_sat=False,
                  max_iterations=1000)
        imp = int(r["records"]["implications"].sum())
        undef = int((r["records"]["status"] == 2).sum())
        procs = 1
    wall = time.perf_counter() - t0
    return {"value": imp / wall, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": f"first {len(sample)} of {len(cubes)} cubes, reference solver as shipped "
                      f"(MAX_ITERATIONS 1000; {undef} cubes ended UNDEF), {imp} implications in {wall:.1f} s"}, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cnf, pre, cubes = make_workload(args.gpus)
    procs = os.cpu_count() or 1
    per_step = min(max(procs * 16, 64), 4096)          # ~16 cubes per core and step: a few seconds per step
    offs, lits = pre.offsets, pre.lits
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        base, wall = cpu_reference_sample(offs, lits, cubes[(i * per_step) % len(cubes):], per_step, procs)
        if i >= args.warmup:
            vals.append(base["value"])
            walls.append(wall)
    value = statistics.mean(vals)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(walls),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.gpus, cubes, 1 if args.gpus > 1 else 0), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, cubes, share=0):
    return {"workload": f"uniform random 3-SAT n={N_VARS} m={N_CLAUSES} r=4.26 splitmix64 seed {SEED} (UNSAT), "
                        f"{len(cubes)} JobChooser cubes of {cubes.shape[1]} literals "
                        f"({len(cubes) // n_gpus} per GPU), full CDCL solve of every cube",
            "cubes": int(len(cubes)), "cube_literals": int(cubes.shape[1]), "parallelism": f"cubes x{n_gpus}",
            "l2": "flushed between timed steps (256 MiB write)", "decision": "vsids", "share_learnts": share}


def c4_sweep(device_index, peak):
    """Config 4 (the HBM-resident configuration): BCP of 1184 jobs x 100 000-literal trails over a planted 3-SAT
    database with n = 1e6, m = 4e6, by the one-CTA-per-job ternary sweep kernel (whole job state in shared memory,
    64-byte bucket index).  Reported beside the headline because it is the one configuration whose clause database
    lives in HBM."""
    import gpupsat_b200 as g
    from gpupsat_b200.instances import planted_3sat_large, sweep_trails
    n, m, J, L = 1_000_000, 4_000_000, 1184, 100_000
    offs, lits, planted = planted_3sat_large(n, m, 4)
    co, cl = sweep_trails(n, J, L, 4, planted)
    with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE, device=device_index) as s:
        s.set_cubes(cube_offsets=co, cube_lits=cl)
        times = []
        for r in range(5):
            got = s.propagate_all(implied_stride=6 * L, want_implied=False)
            if r > 0:
                times.append(s.last_kernel_ms())       # CUDA events on the library's stream around the one launch
        rec = got["records"]
    ms = sum(times) / len(times)
    imp = int(got["n_implied"].sum())
    visited, words = int(rec["watchers_visited"].sum()), int(rec["clause_words_read"].sum())
    alg = 8 * visited + 4 * words + 8 * imp
    traffic, src = None, None
    try:
        import glob
        newest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_sweeptern_ncu_*.json")))[-1]
        d = json.load(open(newest))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        traffic = sum(float(d[k]["value"]) * scale[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        src = f"profiles/{os.path.basename(newest)} (ncu --set full of this launch)"
    except Exception:
        pass
    return {"workload": f"planted 3-SAT n={n} m={m}, {J} jobs x {L}-literal trails, BCP to fixpoint",
            "kernel": "gpsat_bcp_sweep_tern_kernel", "kernel_ms": ms, "kernel_ms_min": min(times),
            "implications_per_s": imp / (ms * 1e-3), "literals_propagated_per_s": (J * L + imp) / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                         "traffic": traffic, "traffic_source": src,
                         "note": "algorithmic bytes = 16 B per occurrence entry visited + 8 B per implication; the bucket "
                                 "index moves 32 or 64 B per literal (a 42-bit entry instead of 16 B), so the DRAM traffic "
                                 "is below the algorithmic bytes; the kernel is bound by shared-memory lookups (LSU ~60 %)"}}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import gpupsat_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = max(world, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    cnf, pre, cubes = make_workload(n_gpus)
    mine = cubes[rank::n_gpus]
    offs, lits = pre.offsets, pre.lits
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    if os.environ.get("GPSAT_WARPS_PER_BLOCK"):
        extra["warps_per_block"] = int(os.environ["GPSAT_WARPS_PER_BLOCK"])
    if os.environ.get("GPSAT_DYNAMIC_SPLIT"):
        extra["dynamic_split"] = int(os.environ["GPSAT_DYNAMIC_SPLIT"])
    for env_name, opt in (("GPSAT_SHARE_LEARNTS", "share_learnts"), ("GPSAT_SPLIT_GAP", "split_gap"),
                          ("GPSAT_SPLIT_BURST", "split_burst"), ("GPSAT_SHARE_MAX_LEN", "share_max_len"),
                          ("GPSAT_SHARE_IMPORT_MAX", "share_import_max"), ("GPSAT_SPLIT_HAND_WORDS", "split_hand_words")):
        if os.environ.get(env_name):
            extra[opt] = int(os.environ[env_name])
    if n_gpus > 1:
        # learnt units and binaries ride the per-epoch all-gather; longer clauses were measured to cost more than they
        # save on this workload (cubes of ~200 conflicts; DESIGN.md "multi-GPU"), GPSAT_SHARE_MAX_LEN overrides
        extra.setdefault("share_learnts", 1)
        extra.setdefault("share_max_len", 2)
    solver = g.Solver(cnf.n_vars, offs, lits, device=local_rank, **extra)
    solver.set_cubes(mine)

    xinfo = {"epochs": 0, "imported_clauses": 0, "exchange_bytes_per_epoch": 0}

    def one_solve():
        """device-resident step; returns (kernel_ms, stats, verdict)"""
        if dist is None:
            verdict, model, st = solver.solve()
            return st["kernel_ms"], st, verdict
        # N > 1: epoch loop of gpupsat_b200.multi_gpu — budgeted persistent-kernel steps, ONE NCCL all-gather per epoch
        # carrying the early-termination flag, the done flags and the short learnt clauses of every rank
        from gpupsat_b200 import multi_gpu as mg
        verdict, model, st, info = mg.solve_sharded(solver, dist, rank, n_gpus, dev, budget_ms=args.epoch_ms,
                                                    max_clauses_per_epoch=1024)
        for k2 in ("epochs", "imported_clauses"):
            xinfo[k2] += info[k2]
        xinfo["exchange_bytes_per_epoch"] = info["exchange_bytes_per_epoch"]
        return st["kernel_ms"], st, verdict

    kernel_ms, stats_acc, launches = [], None, 0
    verdict = None
    with ClockSampler(local_rank) as clocks:
        for i in range(args.warmup):
            one_solve()
            flush.fill_(i & 0xFF)
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            ms, st, verdict = one_solve()
            kernel_ms.append(ms)
            launches += 2 * st["kernel_launches"]                # stamp kernel + solve kernel per launch
            summed = ("jobs_done", "decisions", "implications", "conflicts", "learnt_clauses", "learnt_literals",
                      "restarts", "watchers_visited", "clause_words_read", "kernel_launches", "splits",
                      "warp_busy_frac")
            stats_acc = dict(st) if stats_acc is None else {k: (stats_acc[k] + st[k]) if k in summed else st[k] for k in st}
            flush.fill_(i & 0xFF)                                 # L2 flush between timed steps (outside the event time)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t_wall0)

        # e2e: through the C ABI from host buffers, create -> set_cubes -> solve -> destroy, every step
        e2e_ms, e2e_kernel_ms, e2e_imp = [], [], 0
        h2d = d2h = 0
        for i in range(max(args.steps, 1) + 1):
            barrier()
            t0 = time.perf_counter()
            s2 = g.Solver(cnf.n_vars, offs, lits, device=local_rank, **extra)
            s2.set_cubes(mine)
            v2, m2, st2 = s2.solve()
            s2.close()
            torch.cuda.synchronize()
            if i == 0:
                continue                                          # untimed warm-up of the e2e path (allocator caches)
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            e2e_kernel_ms.append(st2["kernel_ms"])
            e2e_imp += st2["implications"]
        L, m, n = len(lits), len(offs) - 1, cnf.n_vars
        h2d = 4 * ((m + 1) + 2 * (L + m) + (2 * n + 1) + 2 * L + (L + 31) // 32 + 2 * n + (m + 1) + L) + n \
            + 8 * (len(mine) + 1) + 4 * mine.size + 16 + 80 * len(mine)
        d2h = 80 * len(mine) + n + 12 + 8

    steps = args.steps
    tot_ms = sum(kernel_ms)
    imp_local = stats_acc["implications"]
    t = torch.tensor([tot_ms, sum(e2e_ms), sum(e2e_kernel_ms)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([imp_local, e2e_imp, stats_acc["jobs_done"], stats_acc["conflicts"],
                        algorithmic_bytes(stats_acc), stats_acc["decisions"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    tot_ms_max, e2e_ms_max, e2e_kernel_ms_max = float(t[0]), float(t[1]), float(t[2])
    imp_all, e2e_imp_all, jobs_all, confl_all, bytes_all, dec_all = [float(x) for x in cnt]

    if rank == 0:
        peak, peak_src = load_peaks()
        # roofline of the dominant kernel (gpsat_cdcl_kernel) on THIS rank: algorithmic bytes per launch / launch time
        n_launch = max(stats_acc["kernel_launches"], 1)
        bytes_per_launch = algorithmic_bytes(stats_acc) / n_launch
        achieved = (algorithmic_bytes(stats_acc) / (tot_ms * 1e-3)) / 1e9
        cpu_base = None
        c4 = None
        if n_gpus == 1:
            cpu_base, _ = cpu_reference_sample(offs, lits, cubes, 96, 1)
            if os.environ.get("GPSAT_BENCH_C4", "1") != "0":
                try:
                    c4 = c4_sweep(local_rank, peak)
                except Exception as e:          # the headline line must not depend on the secondary measurement
                    c4 = {"error": str(e)[:200]}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_cdcl_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": imp_all / (tot_ms_max * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": tot_ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(n_gpus, cubes, extra.get("share_learnts", 0)),
            "time_to_solve_ms": tot_ms_max / steps, "verdict": {0: "SAT", 1: "UNSAT", 2: "UNDEF"}[verdict],
            "solved_jobs_per_sec": jobs_all / (tot_ms_max * 1e-3), "conflicts_per_sec": confl_all / (tot_ms_max * 1e-3),
            "implications_per_step": imp_all / steps, "wall_ms_timed_region": wall_ms,
            "e2e": {"value": e2e_imp_all / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / max(args.steps, 1),
                    "kernel_ms_per_step": e2e_kernel_ms_max / max(args.steps, 1)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "gpsat_cdcl_kernel",
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "note": "61 KB formula index is staged in shared memory: this kernel is latency / instruction-issue "
                                 "bound and cannot be HBM bound; the HBM fraction is reported because the metric asks for "
                                 "it. c4_sweep is the HBM-resident configuration (DESIGN.md section 7)"},
            "cpu_baseline": cpu_base,
            "c4_sweep": c4,
            "multi_gpu": None if dist is None else {
                "collective": "one NCCL all-gather per epoch (flag + done + short learnt clauses)",
                "epoch_ms": args.epoch_ms, "epochs_per_step": xinfo["epochs"] / max(steps + args.warmup, 1),
                "exchange_bytes_per_epoch": xinfo["exchange_bytes_per_epoch"],
                "clauses_imported_rank0_per_step": xinfo["imported_clauses"] / max(steps + args.warmup, 1)},
            "clocks": clocks.summary(),
            "launch": {"blocks": stats_acc["blocks"], "warps_per_block": stats_acc["warps_per_block"],
                       "smem_bytes_per_block": stats_acc["smem_bytes_per_block"],
                       "state_in_smem": stats_acc["state_in_smem"],
                       "splits_per_step": stats_acc["splits"] / steps,
                       "warp_busy_frac": stats_acc["warp_busy_frac"] / steps},
        }
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--epoch-ms", type=float, default=100.0,
                    help="N > 1: kernel budget per exchange epoch (a C2 shard finishes inside one epoch; measured at "
                         "N=2: 20 ms epochs 62 ms/solve, 60 ms or more 39 ms/solve)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
