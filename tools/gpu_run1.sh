mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q > gpurun_out/r2y_tests.log 2>&1; tail -6 gpurun_out/r2y_tests.log
# ncu: launch list of a short bench run, then one full capture of the CDCL kernel and of the ternary sweep kernel
GPSAT_BENCH_C4=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_a.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2y_ncu_bench.log 2>&1; tail -2 gpurun_out/r2y_ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpsat_cdcl_kernel -s 2 -c 1 -o gpurun_out/r02_cdcl_a python tools/quick_c2.py "" > gpurun_out/r2y_ncu_cdcl.log 2>&1; tail -3 gpurun_out/r2y_ncu_cdcl.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpsat_bcp_sweep_tern -s 2 -c 1 -o gpurun_out/r02_tern_a python tools/sweep_c4.py --jobs 1184 --lens 100000 --reps 1 > gpurun_out/r2y_ncu_tern.log 2>&1; tail -3 gpurun_out/r2y_ncu_tern.log
ls -la gpurun_out/*.ncu-rep | tail -3
