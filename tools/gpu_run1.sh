set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; tail -15 gpurun_out/r2a_tests.log
timeout 300 python tools/quick_c2.py "" > gpurun_out/r2a_quick.log 2>&1; cat gpurun_out/r2a_quick.log
timeout 400 python tools/quick_mesh.py "stride=2" "stride=4" "stride=8" "mesh=2" "mesh=4" "mesh=2 share_learnts=1 share_max_len=2" > gpurun_out/r2a_mesh.log 2>&1; cat gpurun_out/r2a_mesh.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
