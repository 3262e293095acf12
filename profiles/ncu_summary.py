#!/usr/bin/env python
"""Key metrics of one kernel launch from an ncu report as JSON: python profiles/ncu_summary.py X.ncu-rep out.json"""
import csv, json, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__icc_request_hit_rate.pct",
        "sm__icc_requests.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_dshared.sum.pct_of_peak_sustained_elapsed")
d = {}
for i, k in enumerate(hdr):
    if k in want or ("warps_issue_stalled" in k and "per_issue_active" in k):
        try:
            if "stalled" in k and float(vals[i]) < 0.2:
                continue
        except ValueError:
            pass
        d[k] = {"unit": units[i], "value": vals[i]}
try:
    d["git_head"] = {"unit": "", "value": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()}
except Exception:
    pass
if len(sys.argv) > 3:      # extra key=value facts about the profiled launch (e.g. implications_of_this_launch=52670000)
    for kv in sys.argv[3:]:
        k, v = kv.split("=", 1)
        d[k] = {"unit": "", "value": v}
json.dump(d, open(sys.argv[2], "w"), indent=1)
for k, v in d.items():
    print(k, v["value"], v["unit"])
