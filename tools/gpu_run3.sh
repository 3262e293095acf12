mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpsat_cdcl_kernel -s 2 -c 1 -o gpurun_out/r02_cdcl_c python tools/quick_c2.py "" > gpurun_out/r2af_ncu_cdcl.log 2>&1; tail -2 gpurun_out/r2af_ncu_cdcl.log
