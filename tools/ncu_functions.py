#!/usr/bin/env python
"""Dev helper: share of warp instructions and stall samples per function of cdcl_warp.inl from an ncu report captured with
--import-source on.  usage: python tools/ncu_functions.py gpurun_out/X.ncu-rep"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, fname = None, None
agg, samp = collections.Counter(), collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname, hdr = (r[1] if len(r) > 1 else None), None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        ln = int(r[0])
        agg[(fname, ln)] += int(float(r[hdr.index("Instructions Executed")] or 0))
        samp[(fname, ln)] += int(float(r[hdr.index("# Samples")] or 0))
    except Exception:
        pass
src = open(os.path.join(ROOT, "gpupsat_b200/csrc/cdcl_warp.inl")).read().split("\n")
starts = [(i, m.group(1)) for i, l in enumerate(src, 1) for m in [re.match(r"\s*GPSAT_DEV\s+[\w\s\*&]+?\b(\w+)\(", l)] if m]
mid = [i for i, l in enumerate(src, 1) if "(2) learnt clauses watching" in l]
mid = mid[0] if mid else 10 ** 9
def fn(ln):
    name = "?"
    for s_, n in starts:
        if s_ <= ln:
            name = n
    if name == "propagate":
        name = "propagate/originals" if ln < mid else "propagate/learnts"
    return name
fa, fs = collections.Counter(), collections.Counter()
for (f, ln), v in agg.items():
    key = fn(ln) if f and f.endswith("cdcl_warp.inl") else "<" + (f or "?").split("/")[-1] + ">"
    fa[key] += v
    fs[key] += samp[(f, ln)]
ti, ts = sum(fa.values()), sum(fs.values())
print(f"warp instructions {ti}, samples {ts}")
for k, v in fa.most_common(20):
    print(f"{k:28s} inst {100 * v / ti:5.1f}%  samples {100 * fs[k] / max(ts, 1):5.1f}%")
