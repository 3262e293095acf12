mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
