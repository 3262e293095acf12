mkdir -p gpurun_out
REPS=17 timeout 500 python tools/quick_c2.py "" "split_burst=1" "split_burst=2" "split_burst=1,split_gap=12" "split_burst=2,split_gap=12" "split_gap=12" "" "split_burst=1" 2>&1 | tee gpurun_out/quick_c2_split.txt
