#!/usr/bin/env python
"""bench.py — the driver's measurement contract for gpupsat_b200.

Workload (BASELINE.json configs[1]): uniform random 3-SAT n=250 m=1065 (r=4.26, splitmix64 seed 0, UNSAT) split into
4096 JobChooser cubes (reference semantics, -b 8 -t 32 -> k=12).  A step = one complete cube-and-conquer solve of
that instance: every cube is refuted by warp-per-cube CDCL (BCP + 1-UIP learning + VSIDS + geometric restarts).
  value      BCP implications/s over the whole job, formula and cubes already resident in HBM, CUDA-event time of
             the solve kernel on the library's stream (max over ranks)
  e2e        same metric through the C ABI from HOST buffers: gpsat_create (H2D of the formula index) +
             gpsat_set_cubes (H2D) [+ mesh join] + solve (kernel, result reduction, D2H of verdict / per-cube records)
             + gpsat_destroy
  roofline   algorithmic bytes of the solve kernel (8 B per watch entry visited + 4 B per clause word
             read + 8 B per implication, from exact in-kernel counters) / kernel time, against the measured HBM peak;
             `issue` = the bound this kernel can be judged by (warp instructions issued against the SMs' issue slots)
  parity     correctness of the very runs that were timed: verdict, every cube closed, and a per-rank sample of cubes
             re-solved by the CPU oracle (status), plus — N=1 — bit-exact BCP implied lists of all cubes
  cpu_baseline  the reference's own solver classes built for the host (oracle/_ref) on a bounded sample of the same
             cubes (N=1, rank 0 only): one core as shipped, and cap-lifted with a per-cube timeout

N > 1 (torchrun, one process per GPU): STRONG scaling on the SAME job set — the 4096 cubes of N=1 — with the GPUs
joined in a mesh over NVLink peer memory (include/gpsat.h gpsat_mesh_*): a solve is ONE persistent launch per GPU
inside which the cubes are handed out by one cursor that every GPU advances with atomics over NVLink, idle warps take
split-off cubes from the other GPUs' rings, short learnt clauses
are stored into the peers' pools and termination is detected; NCCL carries the IPC handles (one all-gather when the mesh
is formed) and the all-reduces of the per-cube result block at the end.  GPSAT_BENCH_SCALING=weak gives every GPU its
own 4096 cubes (k = 12 + log2 N over the same formula); GPSAT_BENCH_EXCHANGE=nccl runs the epoch loop with one NCCL
all-gather per epoch instead of the mesh (the baseline the mesh is measured against).

`--impl reference` times the reference's CPU implementation (oracle/_ref, all host cores, one process per core over
slices of a bounded cube sample) on the same config / metric; it never loads libgpsat.so.
"""
from __future__ import annotations

import argparse
import json
import os
import platform
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
SMS, SCHEDULERS_PER_SM = 148, 4


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return platform.processor() or "unknown"


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


def scaling_mode():
    return "weak" if os.environ.get("GPSAT_BENCH_SCALING", "strong") == "weak" else "strong"


def cube_request(n_gpus):
    """-b/-t pair handed to the JobChooser: 10*B*T jobs wanted -> k = 12; weak scaling adds log2(N) literals"""
    return (8 * n_gpus if scaling_mode() == "weak" else 8), 32


def make_workload(n_gpus):
    import gpupsat_b200 as g
    from gpupsat_b200.instances import random_ksat
    offs, lits = random_ksat(N_VARS, N_CLAUSES, SEED)
    cnf = g.Cnf.from_arrays(offs, lits)
    pre = cnf.preprocess()
    assert pre.status == g.UNDEF
    b, t = cube_request(n_gpus)
    return cnf, pre, pre.choose_cubes(b, t)


def algorithmic_bytes(stats):
    return 8 * stats["watchers_visited"] + 4 * stats["clause_words_read"] + 8 * stats["implications"]


def workload_config(n_gpus, cubes, share=0, exchange="none"):
    return {"workload": f"uniform random 3-SAT n={N_VARS} m={N_CLAUSES} r=4.26 splitmix64 seed {SEED} (UNSAT), "
                        f"{len(cubes)} JobChooser cubes of {cubes.shape[1]} literals "
                        f"({len(cubes) // n_gpus} per GPU), full CDCL solve of every cube",
            "cubes": int(len(cubes)), "cube_literals": int(cubes.shape[1]), "parallelism": f"cubes x{n_gpus}",
            "l2": "flushed between timed steps (256 MiB write)", "decision": "vsids", "share_learnts": share,
            "multi_gpu": exchange}


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (reference's own classes on host cores).  The only place besides tests/ that executes oracle/.
# ---------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    offs, lits, cubes, nocap = args
    from oracle.binding import Reference
    from oracle.binding import Quiet
    with Quiet():
        R = Reference(offs, lits, nocap=nocap)
        R.reset_counters()
        t = time.perf_counter()
        undef = 0
        for c in cubes:
            st, _ = R.solve(c)
            undef += st == 2
        dt = time.perf_counter() - t
        imp, dec = R.counters()
    return imp, dec, dt, undef, len(cubes)


def reference_workload(n_gpus):
    """The same formula and cubes, produced by the REFERENCE's own preprocessing and JobChooser (oracle/_ref) — the
    reference arm never loads the product library.  tests/test_host_layers.py pins that both give the same arrays."""
    from gpupsat_b200.instances import random_ksat         # pure numpy generator, no library behind it
    from oracle.binding import Quiet, Reference
    offs, lits = random_ksat(N_VARS, N_CLAUSES, SEED)
    with Quiet():
        R = Reference(offs, lits)
        pre_offs, pre_lits = R.formula()
        b, t = cube_request(n_gpus)
        cubes = R.cubes(b, t)
    return pre_offs, pre_lits, cubes


def cpu_reference_sample(pre_offs, pre_lits, cubes, n_sample, procs, nocap=False):
    """Reference solver (as shipped: MAX_ITERATIONS 1000 -> UNDEF on hard cubes; nocap: cap lifted) on a sample."""
    from oracle.binding import REF_DIR
    kind = "reference"
    if not os.path.exists(os.path.join(REF_DIR, "libgpsat_ref.so")):
        kind = "port"
    sample = cubes[:n_sample]
    t0 = time.perf_counter()
    if kind == "reference":
        if procs <= 1:
            res = [_ref_worker((pre_offs, pre_lits, sample, nocap))]
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_ref_worker, [(pre_offs, pre_lits, sample[i::procs], nocap) for i in range(procs)])
        imp = sum(r[0] for r in res)
        undef = sum(r[3] for r in res)
    else:
        from oracle.binding import Oracle
        o = Oracle(N_VARS, pre_offs, pre_lits)
        k = sample.shape[1]
        r = o.run(np.arange(0, sample.size + 1, k, dtype=np.int64), sample.reshape(-1), stop_on_sat=False,
                  max_iterations=0 if nocap else 1000)
        imp = int(r["records"]["implications"].sum())
        undef = int((r["records"]["status"] == 2).sum())
        procs = 1
    wall = time.perf_counter() - t0
    how = "cap lifted (MAX_ITERATIONS removed)" if nocap else "as shipped (MAX_ITERATIONS 1000"
    how += "" if nocap else f"; {undef} cubes ended UNDEF)"
    return {"value": imp / wall, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": f"first {len(sample)} of {len(cubes)} cubes, reference solver {how}, "
                      f"{imp} implications in {wall:.1f} s",
            "cpu_model": cpu_model(), "nproc": os.cpu_count()}, wall


def cap_lifted_estimate(pre_offs, pre_lits, cubes, budget_s=12.0):
    """Time-to-solve comparison the as-shipped reference cannot give (it abandons ~2/3 of the cubes): the cap-lifted
    build on one core, cube after cube, until the time budget is spent; extrapolated to all cubes."""
    import multiprocessing as mp
    from oracle.binding import REF_DIR
    if not os.path.exists(os.path.join(REF_DIR, "libgpsat_ref_nocap.so")):
        return None
    ctx = mp.get_context("fork")
    done, spent, imp = 0, 0.0, 0
    stride = max(1, len(cubes) // 64)              # spread the sample over the cube list
    order = list(range(0, len(cubes), stride))
    with ctx.Pool(1) as pool:
        for j in order:
            left = budget_s - spent
            if left <= 0:
                break
            r = pool.apply_async(_ref_worker, ((pre_offs, pre_lits, cubes[j:j + 1], True),))
            try:
                out = r.get(timeout=left)
            except mp.TimeoutError:
                spent = budget_s
                break
            done += 1
            spent += out[2]
            imp += out[0]
        pool.terminate()
    if done == 0:
        return {"cubes_finished": 0, "seconds": spent, "note": "no cube finished inside the budget"}
    per_cube = spent / done
    return {"cubes_finished": done, "seconds": round(spent, 2), "implications": imp,
            "seconds_per_cube_one_core": per_cube,
            "time_to_solve_s_one_core_extrapolated": per_cube * len(cubes),
            "time_to_solve_s_all_cores_extrapolated": per_cube * len(cubes) / max(os.cpu_count() or 1, 1),
            "note": f"cap-lifted reference on 1 core, every {stride}-th cube until {budget_s:.0f} s are spent; "
                    "a lower bound when the budget cut a cube short"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    offs, lits, cubes = reference_workload(args.gpus)
    procs = os.cpu_count() or 1
    per_step = min(max(procs * 16, 64), 4096)          # ~16 cubes per core and step: a few seconds per step
    vals, walls = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base, wall = cpu_reference_sample(offs, lits, cubes[(i * per_step) % len(cubes):], per_step, procs)
        if i >= args.warmup:
            vals.append(base["value"])
            walls.append(wall)
    value = statistics.mean(vals)
    base["value"] = value
    base["value_per_core"] = value / procs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(walls),
            "higher_is_better": True, "scaling": scaling_mode(), "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(args.gpus, cubes), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# parity of the timed runs (outside the timed region)
# ---------------------------------------------------------------------------------------------------------------
def parity_check(g, pre, cubes, mine_idx, records, verdict, n_gpus, rank, local_rank, extra, dist, dev):
    """records: GLOBAL per-cube records of the last timed solve (every rank holds them after the reduction)."""
    import torch
    from oracle.binding import Oracle
    k = cubes.shape[1]
    closed = int((records["status"] == g.UNSAT).sum())
    ok = verdict == g.UNSAT and closed == len(cubes) and len(records) == len(cubes)
    # a per-rank sample of this rank's own cubes, re-solved by the CPU oracle: the status must agree
    o = Oracle(pre.n_vars, pre.offsets, pre.lits)
    pick = mine_idx[:: max(1, len(mine_idx) // 16)][:16]
    sample = cubes[pick]
    want = o.run(np.arange(0, sample.size + 1, k, dtype=np.int64), sample.reshape(-1), stop_on_sat=False)
    agree = int((want["records"]["status"] == records["status"][pick]).sum())
    ok = ok and agree == len(pick)
    bcp = None
    if n_gpus == 1:
        # the BCP of every cube from the same trail: status and implied literal LISTS bit-exact against the oracle
        with g.Solver(pre.n_vars, pre.offsets, pre.lits, device=local_rank, dynamic_split=0, stop_on_sat=0) as s:
            s.set_cubes(cubes)
            got = s.propagate_all()
        w = o.run(np.arange(0, cubes.size + 1, k, dtype=np.int64), cubes.reshape(-1), mode=1)
        bcp_ok = bool(np.array_equal(got["status"], w["records"]["status"]) and np.array_equal(got["implied"], w["implied"])
                      and np.array_equal(got["conflict_clause"], w["conflict_clause"]))
        bcp = {"cubes": int(len(cubes)), "implied_lists_bit_exact": bcp_ok}
        ok = ok and bcp_ok
    t = torch.tensor([1 if ok else 0, len(pick), agree], dtype=torch.int64, device=dev)
    if dist is not None:
        mn = t.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok_all = bool(int(mn[0]) == 1)
    else:
        ok_all = ok
    return {"verdict": {0: "SAT", 1: "UNSAT", 2: "UNDEF"}[verdict], "cubes": int(len(cubes)), "cubes_closed": closed,
            "sample_checked": int(t[1]), "sample_agree": int(t[2]),
            "sample": "16 cubes per rank re-solved by the CPU oracle (status)", "bcp": bcp, "ok": ok_all}


def c4_sweep(device_index, peak):
    """Config 4 (the HBM-resident configuration): BCP of 1184 jobs x 100 000-literal trails over a planted 3-SAT
    database with n = 1e6, m = 4e6, by the one-CTA-per-job ternary sweep kernel (whole job state in shared memory,
    64-byte bucket index).  Reported beside the headline because it is the one configuration whose clause database
    lives in HBM.  Deviation from BASELINE's wording ("under-constrained random", random trails): the database is
    PLANTED and the trails are drawn from the planted model, because random 10 %-trails over a random database
    conflict within a few hundred literals and the sweep would measure nothing; `random_trails` reports that case too."""
    import gpupsat_b200 as g
    from gpupsat_b200.instances import planted_3sat_large, sweep_trails
    from oracle.binding import Oracle
    n, m, J, L = 1_000_000, 4_000_000, 1184, 100_000
    offs, lits, planted = planted_3sat_large(n, m, 4)
    co, cl = sweep_trails(n, J, L, 4, planted)
    with g.Solver(n, offs, lits, bcp=g.binding.BCP_OCCURRENCE, device=device_index) as s:
        s.set_cubes(cube_offsets=co, cube_lits=cl)
        times = []
        for r in range(5):
            got = s.propagate_all(implied_stride=6 * L, want_implied=(r == 4))
            if r > 0:
                times.append(s.last_kernel_ms())       # CUDA events on the library's stream around the one launch
        rec = got["records"]
        # parity of this very launch: every status, and the implied SETS of 32 jobs against the CPU oracle
        o = Oracle(n, offs, lits)
        pick = list(range(0, J, J // 32))[:32]
        sets_ok, status_ok = True, True
        for j in pick:
            w = o.run(co[j: j + 2] - co[j], cl[co[j]: co[j + 1]], mode=2, implied_stride=6 * L)
            status_ok &= int(w["records"]["status"][0]) == int(got["status"][j])
            sets_ok &= set(got["implied"][j, : got["n_implied"][j]].tolist()) == \
                set(w["implied"][0, : w["n_implied"][0]].tolist())
        all_status = bool((got["status"] == g.UNDEF).all())   # planted trails never conflict
        # non-planted point: random trails over the same database (conflicts included), status vs the oracle
        co2, cl2 = sweep_trails(n, 148, L, 5, None)
        s.set_cubes(cube_offsets=co2, cube_lits=cl2)
        rnd_times = []
        for r in range(3):
            got2 = s.propagate_all(implied_stride=6 * L, want_implied=False)
            if r > 0:
                rnd_times.append(s.last_kernel_ms())
        rnd_ok = True
        for j in range(0, 148, 37):
            w = o.run(co2[j: j + 2] - co2[j], cl2[co2[j]: co2[j + 1]], mode=2, implied_stride=6 * L)
            rnd_ok &= int(w["records"]["status"][0]) == int(got2["status"][j])
        rnd = {"jobs": 148, "kernel_ms": sum(rnd_times) / len(rnd_times),
               "conflicting_jobs": int((got2["status"] == g.UNSAT).sum()),
               "implications": int(got2["n_implied"].sum()),
               "status_agrees_with_oracle_on_4_jobs": bool(rnd_ok)}
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
        src = f"profiles/{os.path.basename(newest)} (ncu --set full of this launch shape; git hash in that file)"
    except Exception:
        pass
    return {"workload": f"planted 3-SAT n={n} m={m}, {J} jobs x {L}-literal trails drawn from the planted model, BCP to fixpoint",
            "kernel": "gpsat_bcp_sweep_tern_kernel", "kernel_ms": ms, "kernel_ms_min": min(times),
            "implications_per_s": imp / (ms * 1e-3), "literals_propagated_per_s": (J * L + imp) / (ms * 1e-3),
            "parity": {"jobs": J, "all_status_fixpoint": all_status, "sets_checked": len(pick),
                       "implied_sets_equal_oracle": bool(sets_ok), "status_equal_oracle": bool(status_ok),
                       "ok": bool(all_status and sets_ok and status_ok)},
            "random_trails": rnd,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                         "frac_closed_form": (80 * (J * L + imp) + 8 * imp) / (ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic, "traffic_source": src,
                         "note": "algorithmic bytes = 16 B per occurrence entry visited + 8 B per implication (counted); "
                                 "frac_closed_form uses SURVEY.md 8d's two-watch model (80 B per propagated literal + 8 B per "
                                 "implication); the bucket index moves 32 or 64 B per literal (a 42-bit entry instead of 16 B), "
                                 "so the DRAM traffic is below the algorithmic bytes; bound by shared-memory lookups"}}


def native_reference_table():
    """Extra (GPSAT_BENCH_NATIVE=1): config 1 in full on the SAME GPU — the reference built natively for sm_100a
    (oracle/_ref/gpupsat_ref_native, test infrastructure) vs the drop-in CLI, -b 1 -t 1, "Total time on GPU" of each."""
    import tempfile
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "gpupsat_ref_native")
    if not os.path.exists(ref_bin):
        return {"unavailable": "oracle/_ref/gpupsat_ref_native not built (make -C oracle native, needs /root/reference)"}
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_native_reference.py"), "--full", "--timeout", "10",
                        "--out", f.name], capture_output=True, text=True, timeout=900)
        rows = json.load(open(f.name))["rows"]
    both = [r for r in rows if r["reference_native"]["gpu_ms"] and r["gpupsat_b200"]["gpu_ms"]]
    return {"instances": len(rows), "reference_finished": len(both),
            "reference_timeouts_10s": sum(r["reference_native"]["verdict"] == "TIMEOUT" for r in rows),
            "verdicts_agree_where_finished": all(r["reference_native"]["verdict"] == r["gpupsat_b200"]["verdict"] for r in both),
            "median_gpu_ms_reference": statistics.median(r["reference_native"]["gpu_ms"] for r in both) if both else None,
            "median_gpu_ms_gpupsat_b200": statistics.median(r["gpupsat_b200"]["gpu_ms"] for r in both) if both else None,
            "rows": rows}


def issue_bound(imp_per_s, clocks):
    """The bound the C2 kernel CAN be judged by: warp instructions issued against the issue slots of 148 SMs x 4
    schedulers at the SM clock sampled during the run.  Warp instructions per implication come from the newest ncu
    capture of the kernel (profiles/r*_cdcl_ncu_*.json: smsp__inst_executed.sum / implications of that launch)."""
    try:
        import glob
        newest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_cdcl_ncu_*.json")))[-1]
        d = json.load(open(newest))
        inst = float(d["smsp__inst_executed.sum"]["value"])
        imp = float(d.get("implications_of_this_launch", {}).get("value", 51.66e6))
        per_imp = inst / imp
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = SMS * SCHEDULERS_PER_SM * mhz * 1e6
        ach = per_imp * imp_per_s
        return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "Ginst/s", "frac": ach / peak,
                "warp_instructions_per_implication": per_imp,
                "source": f"profiles/{os.path.basename(newest)} (instructions per implication) x live implications/s; "
                          f"peak = {SMS} SMs x {SCHEDULERS_PER_SM} schedulers x {mhz:.0f} MHz (sampled during the run)"}
    except Exception as e:
        return {"bound": "issue", "error": str(e)[:120]}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import gpupsat_b200 as g
    from gpupsat_b200 import multi_gpu as mg

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
    exchange = "none" if n_gpus == 1 else os.environ.get("GPSAT_BENCH_EXCHANGE", "mesh")

    cnf, pre, cubes = make_workload(n_gpus)
    n_roots = len(cubes)
    exchange_mesh = exchange == "mesh"
    # mesh: every rank holds ALL cubes (one root cursor for the whole box); nccl epochs: static shard g mod N
    mine_idx = np.arange(n_roots) if (exchange_mesh or n_gpus == 1) else np.arange(rank, n_roots, n_gpus)
    mine = cubes[mine_idx]
    check_idx = np.arange(rank, n_roots, n_gpus)                 # the cubes this rank re-checks against the oracle
    offs, lits = pre.offsets, pre.lits
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    for env_name, opt in (("GPSAT_WARPS_PER_BLOCK", "warps_per_block"), ("GPSAT_DYNAMIC_SPLIT", "dynamic_split"),
                          ("GPSAT_SHARE_LEARNTS", "share_learnts"), ("GPSAT_SPLIT_GAP", "split_gap"),
                          ("GPSAT_SPLIT_BURST", "split_burst"), ("GPSAT_SHARE_MAX_LEN", "share_max_len"),
                          ("GPSAT_SHARE_IMPORT_MAX", "share_import_max"), ("GPSAT_SPLIT_HAND_WORDS", "split_hand_words")):
        if os.environ.get(env_name):
            extra[opt] = int(os.environ[env_name])
    if n_gpus > 1:
        # learnt units and binaries travel between the GPUs (mesh: pushed over NVLink inside the launch; nccl: the
        # per-epoch all-gather); longer clauses were measured to cost more than they save on this workload
        extra.setdefault("share_learnts", 1)
        extra.setdefault("share_max_len", 2)
    solver = g.Solver(cnf.n_vars, offs, lits, device=local_rank, **extra)
    solver.set_cubes(mine)
    block = mg.mesh_join(solver, dist, rank, n_gpus, dev) if exchange == "mesh" else None

    xinfo = {"epochs": 0, "imported_clauses": 0, "exchange_bytes_per_epoch": 0}
    host_ms = {}

    def run_one(s, blk):
        """one complete solve on resident data; returns (kernel_ms of this rank, global-or-local stats, verdict)"""
        if n_gpus == 1:
            verdict, model, st = s.solve()
            return st["kernel_ms"], st, verdict
        if exchange == "mesh":
            # launches bounded to 2 s (a 14-34 ms solve is one launch): a rank whose peer died ends its step instead of
            # spinning until the driver's timeout, and the parity object then says what happened
            verdict, model, st, info = mg.solve_mesh(s, dist, rank, n_gpus, dev, blk, n_roots, budget_ms=2000.0, max_steps=30)
            for k2, v2 in info["host_ms"].items():
                host_ms[k2] = host_ms.get(k2, 0.0) + v2
            host_ms["solves"] = host_ms.get("solves", 0) + 1
            return st["kernel_ms"], st, verdict
        verdict, model, st, info = mg.solve_sharded(s, dist, rank, n_gpus, dev, budget_ms=args.epoch_ms,
                                                    max_clauses_per_epoch=1024)
        for k2 in ("epochs", "imported_clauses"):
            xinfo[k2] += info[k2]
        xinfo["exchange_bytes_per_epoch"] = info["exchange_bytes_per_epoch"]
        return st["kernel_ms"], st, verdict

    summed = ("jobs_done", "decisions", "implications", "conflicts", "learnt_clauses", "learnt_literals", "restarts",
              "watchers_visited", "clause_words_read", "kernel_launches", "splits", "warp_busy_frac", "steals",
              "foreign_clauses")
    kernel_ms, stats_acc, launches = [], None, 0
    verdict = None
    with ClockSampler(local_rank) as clocks:
        for i in range(args.warmup):
            run_one(solver, block)
            flush.fill_(i & 0xFF)
        barrier()
        host_ms.clear()                                           # host phase times: the timed solves only
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            ms, st, verdict = run_one(solver, block)
            kernel_ms.append(ms)
            launches += 2 * st["kernel_launches"] + 1             # stamp kernel + solve kernel per launch, queue init
            stats_acc = dict(st) if stats_acc is None else {k: (stats_acc[k] + st[k]) if k in summed else st[k] for k in st}
            flush.fill_(i & 0xFF)                                 # L2 flush between timed steps (outside the event time)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t_wall0)
        host_ms_timed = dict(host_ms)
        # records of the last timed solve: global after the mesh reduction, local otherwise
        last_records = solver.job_records(n_roots if exchange == "mesh" else None)

        # e2e: through the C ABI from host buffers, create -> set_cubes [-> mesh join] -> solve -> destroy, every step
        e2e_ms, e2e_kernel_ms, e2e_imp = [], [], 0
        e2e_host = {}
        for i in range(max(args.steps, 1) + 1):
            barrier()
            t0 = time.perf_counter()
            s2 = g.Solver(cnf.n_vars, offs, lits, device=local_rank, **extra)
            t1 = time.perf_counter()
            s2.set_cubes(mine)
            t2 = time.perf_counter()
            b2 = mg.mesh_join(s2, dist, rank, n_gpus, dev) if exchange == "mesh" else None
            t3 = time.perf_counter()
            _ms2, st2, v2 = run_one(s2, b2)
            t4 = time.perf_counter()
            s2.close()
            torch.cuda.synchronize()
            t5 = time.perf_counter()
            if i > 0:
                for k2, v3 in (("create", t1 - t0), ("set_cubes", t2 - t1), ("mesh_join", t3 - t2), ("solve", t4 - t3), ("destroy", t5 - t4)):
                    e2e_host[k2] = e2e_host.get(k2, 0.0) + 1e3 * v3 / max(args.steps, 1)
            if i == 0:
                continue                                          # untimed warm-up of the e2e path (allocator caches)
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            e2e_kernel_ms.append(st2["kernel_ms"])
            e2e_imp += st2["implications"]
        L, m, n = len(lits), len(offs) - 1, cnf.n_vars
        h2d = 4 * ((m + 1) + 2 * (L + m) + (2 * n + 1) + 2 * L + (L + 31) // 32 + 2 * n + (m + 1) + L) + n \
            + 8 * (len(mine) + 1) + 4 * mine.size
        d2h = 80 * n_roots + 8 * n_roots + n + 4 * 160

    steps = args.steps
    tot_ms = sum(kernel_ms)
    mesh_global = exchange == "mesh"                              # stats already summed over ranks by the reduction
    t = torch.tensor([tot_ms, sum(e2e_ms), sum(e2e_kernel_ms)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([stats_acc["implications"], e2e_imp, stats_acc["jobs_done"], stats_acc["conflicts"],
                        algorithmic_bytes(stats_acc), stats_acc["decisions"], stats_acc["splits"]],
                       dtype=torch.float64, device=dev)
    per_rank = torch.tensor([stats_acc["warp_busy_frac"] / steps, stats_acc["steals"] / steps,
                             stats_acc["foreign_clauses"] / steps], dtype=torch.float64, device=dev)
    ranks_busy = [per_rank.clone() for _ in range(n_gpus)]
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if not mesh_global:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_gather(ranks_busy, per_rank)
    else:
        ranks_busy = [per_rank]
    tot_ms_max, e2e_ms_max, e2e_kernel_ms_max = float(t[0]), float(t[1]), float(t[2])
    imp_all, e2e_imp_all, jobs_all, confl_all, bytes_all, dec_all, splits_all = [float(x) for x in cnt]
    if not mesh_global and dist is not None:
        # nccl-epoch mode: records are per rank; gather the statuses for the parity check
        full = torch.zeros(n_roots, dtype=torch.int32, device=dev)
        full[torch.as_tensor(mine_idx, device=dev)] = torch.as_tensor(last_records["status"].astype(np.int32), device=dev)
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        glob_rec = np.zeros(n_roots, dtype=g.RECORD_DTYPE)
        glob_rec["status"] = full.cpu().numpy()
        last_records = glob_rec
    parity = parity_check(g, pre, cubes, check_idx, last_records, verdict, n_gpus, rank, local_rank, extra, dist, dev)

    if rank == 0:
        peak, peak_src = load_peaks()
        clk = clocks.summary()
        # roofline of the dominant kernel (gpsat_cdcl_kernel): algorithmic bytes per launch / launch time, all ranks
        n_launch = max(stats_acc["kernel_launches"], 1)
        achieved = (bytes_all / (tot_ms_max * 1e-3)) / 1e9 / n_gpus          # per GPU
        cpu_base = None
        c4 = None
        if n_gpus == 1:
            cpu_base, _ = cpu_reference_sample(offs, lits, cubes, 96, 1)
            try:
                cpu_base["cap_lifted"] = cap_lifted_estimate(offs, lits, cubes)
            except Exception as e:
                cpu_base["cap_lifted"] = {"error": str(e)[:200]}
            if os.environ.get("GPSAT_BENCH_C4", "1") != "0":
                try:
                    c4 = c4_sweep(local_rank, peak)
                except Exception as e:          # the headline line must not depend on the secondary measurement
                    c4 = {"error": str(e)[:200]}
        traffic = None
        try:
            import glob
            tp = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_cdcl_traffic*.json")))[-1]
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
        value = imp_all / (tot_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": tot_ms_max / steps, "higher_is_better": True,
            "scaling": scaling_mode() if n_gpus > 1 else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(n_gpus, cubes, extra.get("share_learnts", 0), exchange),
            "time_to_solve_ms": tot_ms_max / steps, "verdict": {0: "SAT", 1: "UNSAT", 2: "UNDEF"}[verdict],
            "solved_jobs_per_sec": jobs_all / (tot_ms_max * 1e-3), "conflicts_per_sec": confl_all / (tot_ms_max * 1e-3),
            "implications_per_step": imp_all / steps, "wall_ms_timed_region": wall_ms,
            "parity": parity,
            "e2e": {"value": e2e_imp_all / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / max(args.steps, 1),
                    "kernel_ms_per_step": e2e_kernel_ms_max / max(args.steps, 1),
                    "host_ms_rank0": e2e_host},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "gpsat_cdcl_kernel",
                         "algorithmic_bytes_per_launch": bytes_all / n_launch / n_gpus,
                         "issue": issue_bound(value / n_gpus, clk),
                         "note": "the formula index (31 KB packed) is staged in shared memory: this kernel is instruction-issue / latency "
                                 "bound and cannot be HBM bound; the HBM fraction is reported because the metric asks for "
                                 "it, `issue` is the bound it can be judged by. c4_sweep is the HBM-resident configuration "
                                 "(DESIGN.md section 7)"},
            "cpu_baseline": cpu_base,
            "c4_sweep": c4,
            "native_reference_c1": native_reference_table() if (n_gpus == 1 and os.environ.get("GPSAT_BENCH_NATIVE") == "1") else None,
            "multi_gpu": None if dist is None else {
                "mode": exchange,
                "collective": ("mesh over NVLink peer memory inside ONE launch per GPU (steals, clause push, termination); "
                               "NCCL: all-gather of IPC handles at join, 3 all-reduces of the result block per solve")
                if exchange == "mesh" else "one NCCL all-gather per epoch (flag + done + short learnt clauses)",
                "epoch_ms": None if exchange == "mesh" else args.epoch_ms,
                "epochs_per_step": None if exchange == "mesh" else xinfo["epochs"] / max(steps + args.warmup, 1),
                "exchange_bytes_per_epoch": None if exchange == "mesh" else xinfo["exchange_bytes_per_epoch"],
                "result_block_bytes": 4 * int(block.numel()) if block is not None else None,
                "host_ms_per_solve_rank0": {k2: v2 / max(host_ms_timed.get("solves", 1), 1) for k2, v2 in host_ms_timed.items() if k2 != "solves"} or None,
                "per_rank": [{"warp_busy_frac": float(x[0]), "steals_per_step": float(x[1]),
                              "foreign_clauses_per_step": float(x[2])} for x in ranks_busy]},
            "clocks": clk,
            "launch": {"blocks": stats_acc["blocks"], "warps_per_block": stats_acc["warps_per_block"],
                       "smem_bytes_per_block": stats_acc["smem_bytes_per_block"],
                       "state_in_smem": stats_acc["state_in_smem"],
                       "splits_per_step": splits_all / steps,
                       "warp_busy_frac": float(sum(float(x[0]) for x in ranks_busy) / len(ranks_busy))},
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
                    help="GPSAT_BENCH_EXCHANGE=nccl only: kernel budget per exchange epoch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
