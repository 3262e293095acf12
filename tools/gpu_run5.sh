timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -1
REPS=5 timeout 60 python tools/quick_c2.py "" "warps_per_block=28" 2>&1 | cut -c1-150
