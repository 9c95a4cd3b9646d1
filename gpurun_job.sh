mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 200 --warmup 10 --no-cpu-baseline | cut -c1-150
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --workload c3 | cut -c1-150
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"move_kernel|paint_kernel" -s 10 -c 2 --csv --log-file gpurun_out/k.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
PAINTRL_PROFILE=1 python -m paintrl_b200.build --force > /dev/null 2>&1
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python profiles/phase_profile.py --envs 148 --no-flush 2>&1 | grep -E "^\s+\[(1[5-9]|2[0-9])\]|total"
