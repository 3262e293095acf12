timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -x -q -m gpu -k "bit_exact or dynamic_split or budgeted or mesh_small or golden" 2>&1 | tail -1
REPS=13 timeout 100 python tools/quick_c2.py "" 2>&1 | cut -c1-170
timeout 60 python tools/phase_c2.py 2>&1 | sed -n 1,3p
