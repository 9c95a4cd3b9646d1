mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | cut -c1-150 | tail -1
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --workload c3 2>&1 | cut -c1-150 | tail -1
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=8 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --workload c5 | cut -c1-150
