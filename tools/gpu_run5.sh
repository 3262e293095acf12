mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_exact or dynamic_split or golden" 2>&1 | tail -2
REPS=13 timeout 500 python tools/quick_c2.py "" "warps_per_block=28" "warps_per_block=20" "" 2>&1 | tee gpurun_out/quick_c2_uniform.txt
