mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; head -c 330 gpurun_out/r02_bench_f.json; echo; tail -1 gpurun_out/r02_bench_f.err
