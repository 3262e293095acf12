mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_scale4_e.json 2> gpurun_out/r02_scale4_e.err; tail -1 gpurun_out/r02_scale4_e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_scale4_e.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["launch"])
PY
