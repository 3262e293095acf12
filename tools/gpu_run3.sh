mkdir -p gpurun_out
timeout 600 python tools/quick_mesh.py "" "split_hard=64" "split_hard=128" "split_hard=256" "split_hard=128 split_burst=8" "stride=8" "stride=8 split_hard=64" "stride=8 split_hard=128" "stride=8 split_hard=256" "stride=8 split_hard=128 split_burst=8" 2>&1 | grep -v children > gpurun_out/r2u_mesh.log; cat gpurun_out/r2u_mesh.log
