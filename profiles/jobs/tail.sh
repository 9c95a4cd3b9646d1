python -m pytest tests/test_gpu_variants.py tests/test_gpu_baseline_sizes.py tests/test_gpu_oracle_batch.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
PAINTRL_FUSED=1 python profiles/step_tail.py 2>&1 | tail -7
PAINTRL_FUSED=0 python profiles/step_tail.py 2>&1 | tail -7
PAINTRL_FUSED=1 python profiles/step_tail.py --workload c4 --envs 2048 --steps 60 2>&1 | tail -5
