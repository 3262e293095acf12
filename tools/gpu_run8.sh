timeout 300 python tools/quick_mesh.py "gpus=8 split_reserve=-1" "gpus=8" "gpus=8 split_reserve=128" "gpus=8 split_reserve=8" "gpus=4" "gpus=2" "" 2>&1 | grep -v children
