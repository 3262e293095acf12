mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_final.log 2>&1; tail -4 gpurun_out/r02_tests_final.log
