mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
REPS=13 timeout 300 python tools/quick_c2.py "" "phase_stats=1" "warps_per_block=28" "" 2>&1 | tee gpurun_out/quick_c2_lean2.txt
