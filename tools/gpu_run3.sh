mkdir -p gpurun_out
timeout 600 python tools/quick_mesh.py "" "stride=8" "mesh=2" "blocks=74" 2>&1 | grep -v children > gpurun_out/r2z_mesh.log; cat gpurun_out/r2z_mesh.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -x -q 2>&1 | tail -3
