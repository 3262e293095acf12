timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2ai_bench8.json 2> gpurun_out/r2ai_bench8.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ai_bench8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["parity"]["ok"], d["e2e"], d["multi_gpu"]["host_ms_per_solve_rank0"])
PY
tail -2 gpurun_out/r2ai_bench8.err
