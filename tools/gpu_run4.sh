mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "layouts or config4" 2>&1 | tail -2
timeout 900 python tools/sweep_c4.py --jobs 1184 --lens 100000,200000 --flags ${FLAGS:-0,64,2} --out gpurun_out/r02_c4_variants.json > gpurun_out/r02_c4_variants.log 2>&1
cut -c1-90 gpurun_out/r02_c4_variants.log | tail -10
