#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python profiles/ncu_lines.py gpurun_out/X.ncu-rep [top]
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = {}
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
            lsb, ssb, wt = hdr.index("stall_long_sb"), hdr.index("stall_short_sb"), hdr.index("stall_wait")
            continue
        if hdr is None or r[0] in ("", "Function Name") or len(r) <= wt:
            continue
        try:
            s, n = int(r[si]), int(r[ie])
            l, sh, w = int(r[lsb]), int(r[ssb]), int(r[wt])
        except ValueError:
            continue
        key = (fname, int(r[0]), r[1].strip())
        a = agg.setdefault(key, [0, 0, 0, 0, 0])
        for i, v in enumerate((s, n, l, sh, w)):
            a[i] += v
    tot = sum(v[0] for v in agg.values()) or 1
    totn = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot}, warp instructions {totn}")
    print("  smp%  inst%  long_sb short_sb  wait  file:line  source")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:6.2f} {100 * v[1] / totn:6.2f} {v[2]:8d} {v[3]:8d} {v[4]:6d}  {k[0]}:{k[1]}  {k[2][:100]}")


if __name__ == "__main__":
    main()
