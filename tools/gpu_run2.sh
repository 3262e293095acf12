GPSAT_DEBUG_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2aj_bench2.json 2> gpurun_out/r2aj_bench2.err; grep "begin:" gpurun_out/r2aj_bench2.err | tail -8
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2aj_bench2.json").read().strip().splitlines()[-1])
print(d["e2e"], d["multi_gpu"]["host_ms_per_solve_rank0"])
PY
