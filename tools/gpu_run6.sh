mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; head -c 400 gpurun_out/r02_bench_e.json; echo; tail -2 gpurun_out/r02_bench_e.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_e.json 2>/dev/null
GPSAT_BENCH_C4=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_e.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_ncu_bench.log 2>&1
timeout 900 python tools/sweep_c4.py --out gpurun_out/r02_c4_sweep_e.json > gpurun_out/r2e_sweep.log 2>&1; tail -9 gpurun_out/r2e_sweep.log | cut -c1-100
timeout 300 python tools/timeline.py 1 > gpurun_out/r02_timeline_c2_e.txt 2>&1; head -3 gpurun_out/r02_timeline_c2_e.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gpsat_bcp_sweep_tern -s 1 -c 1 -f -o gpurun_out/r02_sweeptern_e python tools/sweep_c4.py --jobs 1184 --lens 100000 --reps 2 > gpurun_out/r2e_ncu_tern.log 2>&1; tail -2 gpurun_out/r2e_ncu_tern.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gpsat_cdcl_kernel -s 3 -c 1 -f -o gpurun_out/r02_cdcl_d python tools/quick_c2.py "" > gpurun_out/r2e_ncu_cdcl.log 2>&1; tail -2 gpurun_out/r2e_ncu_cdcl.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
