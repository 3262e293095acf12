set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_multi.py -x -q > gpurun_out/r2b_tests.log 2>&1; tail -15 gpurun_out/r2b_tests.log
timeout 600 python tools/quick_mesh.py "" "blocks=74" "stride=8" "stride=8 split_at_start=-1" "stride=8 split_gap_hot=1 split_burst=8" "mesh=2" "mesh=2 mesh_flags=1" "gpus=2" "gpus=2 mesh_flags=1" "gpus=2 share_learnts=1 share_max_len=2" > gpurun_out/r2b_mesh.log 2>&1; cat gpurun_out/r2b_mesh.log
GPSAT_BENCH_C4=0 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; head -c 600 gpurun_out/r2b_bench1.json; tail -3 gpurun_out/r2b_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; cat gpurun_out/r2b_bench2.json; tail -5 gpurun_out/r2b_bench2.err
GPSAT_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench2n.json 2> gpurun_out/r2b_bench2n.err; head -c 700 gpurun_out/r2b_bench2n.json; tail -5 gpurun_out/r2b_bench2n.err
