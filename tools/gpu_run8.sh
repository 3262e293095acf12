mkdir -p gpurun_out
timeout 900 python tools/run_configs_multi.py c5 --out gpurun_out/r2x_c5.json > gpurun_out/r2x_c5.log 2>&1; python - <<'PY'
import json
for r in json.load(open("gpurun_out/r2x_c5.json"))["rows"]:
    print("c5", r["gpus"], "share", r["share_max_len"], r["verdict"], "closed", r["cubes_closed"], "ms", round(r["kernel_ms"],2), "confl", r["conflicts"], "recv", r["clauses_received_over_nvlink"], "steals", r["steals"], "busy", round(r["warp_busy_frac"],2))
PY
tail -3 gpurun_out/r2x_c5.log
timeout 900 python tools/run_configs_multi.py c3 --gpus 2 4 8 --seconds 6 --out gpurun_out/r2x_c3.json > gpurun_out/r2x_c3.log 2>&1; python - <<'PY'
import json
for r in json.load(open("gpurun_out/r2x_c3.json"))["rows"]:
    print("c3", r["gpus"], "share", r["share_max_len"], r["verdict"], "closed", r["cubes_closed"], "closed/s", round(r["cubes_closed_per_s"],1), "confl/s", f'{r["conflicts_per_s"]:.3e}', "impl/s", f'{r["implications_per_s"]:.3e}', "recv", r["clauses_received_over_nvlink"], "wall", round(r["wall_s"],2), "busy", round(r["warp_busy_frac"],2))
PY
tail -3 gpurun_out/r2x_c3.log
