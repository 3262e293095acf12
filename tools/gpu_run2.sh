mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 600 python tools/quick_mesh.py "" "gpus=2" "gpus=2 mesh_flags=1" "gpus=2 share_learnts=1 share_max_len=2" "gpus=2 stride=8" 2>&1 | grep -v children > gpurun_out/r2t_mesh.log; cat gpurun_out/r2t_mesh.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2t_bench2.json 2> gpurun_out/r2t_bench2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","implications_per_step","parity")}, d["e2e"], d["multi_gpu"]["per_rank"], d["launch"])
PY
tail -3 gpurun_out/r2t_bench2.err
