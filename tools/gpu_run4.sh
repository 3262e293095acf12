mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "occurrence or config4" 2>&1 | tail -2
timeout 900 python tools/sweep_c4.py --jobs 1184 --lens 100000 --flags ${FLAGS:-1,17} --out gpurun_out/r02_c4_general.json > gpurun_out/r02_c4_general.log 2>&1
cut -c1-100 gpurun_out/r02_c4_general.log | tail -6
