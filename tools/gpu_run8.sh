mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python tools/quick_mesh.py "" "gpus=2" "gpus=4" "gpus=8" "gpus=8 mesh_flags=1" "gpus=8 split_gap=4" "gpus=8 split_hard=64" 2>&1 | grep -v children > gpurun_out/r2v_mesh.log; cat gpurun_out/r2v_mesh.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2v_bench$n.json 2> gpurun_out/r2v_bench$n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2v_bench$n.json").read().strip().splitlines()[-1])
print($n, {k:d[k] for k in ("value","ms_per_step","implications_per_step","parity")}, d["e2e"], [round(x["warp_busy_frac"],2) for x in d["multi_gpu"]["per_rank"]], [int(x["steals_per_step"]) for x in d["multi_gpu"]["per_rank"]], d["launch"])
PY
tail -2 gpurun_out/r2v_bench$n.err
done
