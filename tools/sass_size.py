#!/usr/bin/env python
"""Dev helper: SASS instruction count of gpsat_cdcl_kernel<true,true> per source function of cdcl_warp.inl
(instruction-cache footprint is a first-order cost of this kernel).  usage: python tools/sass_size.py"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "gpupsat_b200/csrc/cdcl_warp.inl")).read().split("\n")
# function start lines: "GPSAT_DEV <type> name("
starts = []
for i, ln in enumerate(src, 1):
    m = re.match(r"\s*GPSAT_DEV\s+[\w\s\*&]+?\b(\w+)\(", ln)
    if m:
        starts.append((i, m.group(1)))
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "gpupsat_b200/libgpsat.so")], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith("kernels")][0]
    out = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, kern = None, None
cnt = collections.Counter()
for line in out.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\.text\.(\S+):", line)
    if m:
        kern = m.group(1)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line) and kern and "Lb1ELb1" in kern:
        cnt[cur] += 1
agg = collections.Counter()
for (f, l), c in cnt.items():
    if f == "cdcl_warp.inl":
        name = "?"
        for s, n in starts:
            if s <= l:
                name = n
        agg[name] += c
    else:
        agg["<" + f + ">"] += c
print("total", sum(cnt.values()))
for n, c in agg.most_common():
    print(f"{c:6d}  {n}")
