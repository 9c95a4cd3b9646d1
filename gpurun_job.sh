mkdir -p gpurun_out
PAINTRL_PROFILE=1 python -m paintrl_b200.build --force > /dev/null 2>&1
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python profiles/phase_profile.py --envs 4096 --no-flush 2>&1 | tee gpurun_out/phase_c2_warm.txt
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python profiles/phase_profile.py --envs 148 --no-flush 2>&1 | tee gpurun_out/phase_c2_148_warm.txt
python -m paintrl_b200.build --force > /dev/null 2>&1
PAINTRL_MOVE_MINB=4 PAINTRL_MOVE_LANES=32 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-flush | cut -c1-200
