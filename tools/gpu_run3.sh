mkdir -p gpurun_out
timeout 900 python tools/run_native_reference.py --full --timeout 10 --out gpurun_out/r02_native_reference.json > gpurun_out/r2ac_native.log 2>&1; tail -3 gpurun_out/r2ac_native.log | cut -c1-250
for u in 0 3 4; do
  touch gpupsat_b200/csrc/kernels.cu
  if [ $u = 0 ]; then make lib > /dev/null 2>&1; else make lib EXTRA_NVFLAGS=-DGPSAT_HOTLOOP_UNROLL=$u > /dev/null 2>&1; fi
  echo "== hot loop unroll $u"; timeout 300 python tools/quick_c2.py "" "" 2>&1 | tail -2
done
