timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "longer_than or single_cube" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ah_bench2.json 2> gpurun_out/r2ah_bench2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ah_bench2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","implications_per_step")}, d["parity"]["ok"], d["e2e"]["ms_per_step"])
PY
tail -2 gpurun_out/r2ah_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-200
