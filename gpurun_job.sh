mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | cut -c1-150
