mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; tail -12 gpurun_out/r2w_tests.log
