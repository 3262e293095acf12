mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; head -c 900 gpurun_out/r02_bench_b.json; echo; tail -2 gpurun_out/r02_bench_b.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_b.json 2>/dev/null; head -c 400 gpurun_out/r02_bench_ref_b.json; echo
GPSAT_BENCH_C4=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2aa_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpsat_cdcl_kernel -s 2 -c 1 -o gpurun_out/r02_cdcl_b python tools/quick_c2.py "" > gpurun_out/r2aa_ncu_cdcl.log 2>&1; tail -2 gpurun_out/r2aa_ncu_cdcl.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpsat_bcp_sweep_tern -s 1 -c 1 -o gpurun_out/r02_tern_b python tools/sweep_c4.py --jobs 1184 --lens 100000 --reps 2 > gpurun_out/r2aa_ncu_tern.log 2>&1; tail -2 gpurun_out/r2aa_ncu_tern.log
timeout 900 python tools/sweep_c4.py --out gpurun_out/r02_c4_sweep_b.json > gpurun_out/r2aa_sweep.log 2>&1; tail -9 gpurun_out/r2aa_sweep.log | cut -c1-200
