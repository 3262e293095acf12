mkdir -p gpurun_out
timeout 300 python tools/quick_mesh.py "" "gpus=2" "gpus=4" "gpus=8" 2>&1 | grep -v children > gpurun_out/r2ab_mesh.log; cat gpurun_out/r2ab_mesh.log
GPSAT_BENCH_C4=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_scale1_c.json 2>/dev/null
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_scale${n}_c.json 2> gpurun_out/r02_scale${n}_c.err
done
for n in 1 2 4 8; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02_scale${n}_c.json").read().strip().splitlines()[-1])
print($n, "value %.3e" % d["value"], "ms %.2f" % d["ms_per_step"], "impl/step %.3e" % d["implications_per_step"], "parity", d["parity"]["ok"], d["parity"]["cubes_closed"], "e2e ms %.2f" % d["e2e"]["ms_per_step"], "busy %.2f" % d["launch"]["warp_busy_frac"], "splits %d" % d["launch"]["splits_per_step"])
PY
done
