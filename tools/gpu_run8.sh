mkdir -p gpurun_out
timeout 300 python tools/quick_mesh.py "" "gpus=2" "gpus=4" "gpus=8" "gpus=8 mesh_flags=1" 2>&1 | grep -v children > gpurun_out/r02_mesh_scaling_quick_d.txt; cat gpurun_out/r02_mesh_scaling_quick_d.txt
GPSAT_BENCH_C4=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_scale1_d.json 2>/dev/null
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2980$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale${n}_d.json 2> gpurun_out/r02_scale${n}_d.err
done
GPSAT_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29899 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_scale8_nccl_d.json 2>/dev/null
for n in 1 2 4 8 8_nccl; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02_scale${n}_d.json").read().strip().splitlines()[-1])
print("$n", "value %.3e" % d["value"], "ms %.2f" % d["ms_per_step"], "impl/step %.3e" % d["implications_per_step"], "parity", d["parity"]["ok"], d["parity"]["cubes_closed"], "e2e ms %.2f" % d["e2e"]["ms_per_step"], "busy %.2f" % d["launch"]["warp_busy_frac"], "splits %d" % d["launch"]["splits_per_step"])
PY
done
timeout 600 python tools/run_configs_multi.py c5 --gpus 1 8 --out gpurun_out/r02_config5_d.json > /dev/null 2>&1; python - <<'PY'
import json
for r in json.load(open("gpurun_out/r02_config5_d.json"))["rows"]:
    print("c5", r["gpus"], "share", r["share_max_len"], r["verdict"], "ms", round(r["kernel_ms"],2), "confl", r["conflicts"])
PY
