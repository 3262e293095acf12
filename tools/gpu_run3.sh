mkdir -p gpurun_out
for ns in 4000 1000; do
make lib EXTRA_NVFLAGS=-DGPSAT_IDLE_NS=${ns}u > /dev/null 2>&1 || echo BUILD FAILED
echo "== idle base ${ns} ns"
timeout 600 python tools/quick_mesh.py "" "split_gap=4" "split_gap_hot=2" "split_gap_hot=4" "split_gap=4 split_gap_hot=2" "split_gap=4 split_gap_hot=1 split_burst=2" "stride=8" "stride=8 split_gap=4" "stride=8 split_gap_hot=2" "stride=8 split_gap_hot=4"  "stride=8 split_gap=4 split_gap_hot=2"  "stride=8 split_gap=4 split_gap_hot=1 split_burst=2" "stride=8 split_gap=4 split_gap_hot=2 split_burst=8" 2>&1 | grep -v children
done > gpurun_out/r2i_mesh.log 2>&1; cat gpurun_out/r2i_mesh.log
