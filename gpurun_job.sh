mkdir -p gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_oracle_batch.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | tee gpurun_out/bench_c2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step']*1e3, d['e2e'])"
python bench.py --workload c5 --rollout --steps 3 --warmup 3 | tee gpurun_out/bench_c5_rollout.json | cut -c1-1200
python bench.py --workload c5 --steps 50 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_c5.json | cut -c1-300
