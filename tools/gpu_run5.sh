mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/quick_c2.py "warps_per_block=24" "" "warps_per_block=26" "warps_per_block=20" 2>&1 | tee gpurun_out/quick_c2.txt
