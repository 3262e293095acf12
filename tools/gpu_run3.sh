timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python tools/quick_mesh.py "" "" "stride=8" 2>&1 | grep -v children
