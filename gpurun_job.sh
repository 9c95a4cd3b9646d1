mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py --workload c5 --rollout --steps 3 --warmup 3 | tee gpurun_out/bench_c5_rollout.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step']*1e3, d['config']['workload'][-80:])"
python bench.py --workload c2 --rollout --steps 5 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 rollout', d['value']/1e6, d['ms_per_step']*1e3, d['config']['workload'][-80:])"
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6)"
