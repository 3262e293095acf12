#!/usr/bin/env python
"""Dev helper: run bench.py under several GPSAT_* environment settings and print one short line per run.
usage: python tools/bench_sweep.py "GPSAT_DYNAMIC_SPLIT=0" "GPSAT_DYNAMIC_SPLIT=1 GPSAT_SHARE_LEARNTS=1" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=")
        env[k] = v
    steps = env.get("SWEEP_STEPS", "2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "1"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"[{spec}] warps/block {d['launch']['warps_per_block']} blocks {d['launch']['blocks']} "
              f"ms/step {d['ms_per_step']:.2f} impl/s {d['value']:.3e} impl/step {d['implications_per_step']:.3e} "
              f"confl/s {d['conflicts_per_sec']:.3e} busy {d['launch']['warp_busy_frac']:.2f} splits {d['launch']['splits_per_step']:.0f} e2e_ms {d['e2e']['ms_per_step']:.2f} {d['verdict']}", flush=True)
    except Exception as e:
        print(f"[{spec}] FAILED {e}: {out.stdout[-500:]} {out.stderr[-1500:]}", flush=True)
