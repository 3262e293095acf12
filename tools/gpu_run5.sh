mkdir -p gpurun_out
timeout 300 python tools/run_configs_multi.py c5 --gpus 1 --out gpurun_out/r02_config5_php_10_9_e.json 2>&1 | cut -c1-260 | tail -6
