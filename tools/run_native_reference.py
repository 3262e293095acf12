#!/usr/bin/env python
"""Same-silicon baseline: the UNMODIFIED reference built natively for sm_100a
(oracle/ref/native/build_native.sh -> oracle/_ref/gpupsat_ref_native; test infrastructure, Boost-free front end) and this
repo's drop-in CLI on the same small instances (config 1: uf20-91, uf50-218), sequential (-b 1 -t 1) and, where the
reference's 100 000-node pool allows it, parallel.  Prints one JSON object: per instance and mode the verdict and
"Total time on GPU" of both binaries.  usage: python tools/run_native_reference.py [--out FILE]"""
import argparse, json, os, re, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpupsat_b200.instances import random_ksat, to_dimacs  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "gpupsat_ref_native")
OURS = os.path.join(ROOT, "gpupsat_b200", "gpupsat")


TIMEOUT_S = 20.0


def run(binary, path, mode, cwd):
    t = time.perf_counter()
    try:
        out = subprocess.run([binary, path] + mode, capture_output=True, text=True, cwd=cwd, timeout=TIMEOUT_S)
        text = out.stdout
        rc = out.returncode
    except subprocess.TimeoutExpired:
        return {"verdict": "TIMEOUT", "gpu_ms": None, "wall_s": TIMEOUT_S, "rc": None}
    wall = time.perf_counter() - t
    m = re.search(r"Total time on GPU: ([0-9.]+) ms", text)
    verdict = next((v for v in ("UNSATISFIABLE", "SATISFIABLE", "UNDEFINED") if re.search(rf"^{v}$", text, re.M)), "?")
    return {"verdict": verdict, "gpu_ms": float(m.group(1)) if m else None, "wall_s": round(wall, 3), "rc": rc}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--full", action="store_true", help="config 1 in full: the reference's 6 tests/cnf files, uf20-91 and "
                    "uf50-218 seeds 0-9, sequential mode (-b 1 -t 1)")
    ap.add_argument("--timeout", type=float, default=20.0)
    args = ap.parse_args()
    global TIMEOUT_S
    TIMEOUT_S = args.timeout
    rows = []
    with tempfile.TemporaryDirectory() as d:
        cases = []
        if args.full:
            G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")))
            for name, rec in sorted(G["tests_cnf"].items()):
                path = os.path.join(d, name)
                open(path, "w").write(rec["dimacs"])
                cases.append((path, [["-b", "1", "-t", "1"]]))
            seeds = range(10)
        else:
            seeds = (0, 1)
        for n, m in ((20, 91), (50, 218)):
            for seed in seeds:
                offs, lits = random_ksat(n, m, seed)
                path = os.path.join(d, f"uf{n}-{m}-{seed}.cnf")
                open(path, "w").write(to_dimacs(offs, lits, n))
                # parallel mode only where 2 * m * B * T nodes fit the reference's pool of 100 000
                modes = [["-b", "1", "-t", "1"]] + ([["-b", "4", "-t", "32"]] if (2 * m * 128 <= 100000 and not args.full) else [])
                cases.append((path, modes))
        for path, modes in cases:
            if True:
                for mode in modes:
                    row = {"instance": os.path.basename(path), "mode": " ".join(mode),
                           "reference_native": run(REF, path, mode, d), "gpupsat_b200": run(OURS, path, mode, d)}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                    if args.out:      # rewritten after every row: a run cut short keeps what it measured
                        json.dump({"what": "reference (native sm_100a build, unmodified kernels) vs gpupsat_b200 CLI on "
                                           "one B200", "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
