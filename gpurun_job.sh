mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
PAINTRL_MOVE_LANES=32 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PAINTRL_DEBUG=1 PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | cut -c1-150 | tail -3
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:"move_kernel|paint_kernel" -s 10 -c 2 --csv --log-file gpurun_out/k.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
