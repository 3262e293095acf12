mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; head -c 300 gpurun_out/r02_bench_d.json; echo; tail -2 gpurun_out/r02_bench_d.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_d.json 2>/dev/null
GPSAT_BENCH_C4=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_d.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2ag_ncu_bench.log 2>&1
timeout 900 python tools/sweep_c4.py --out gpurun_out/r02_c4_sweep_d.json > gpurun_out/r2ag_sweep.log 2>&1; tail -9 gpurun_out/r2ag_sweep.log | cut -c1-60,300-420
timeout 300 python tools/timeline.py 1 > gpurun_out/r02_timeline_c2_d.txt 2>&1; head -3 gpurun_out/r02_timeline_c2_d.txt
