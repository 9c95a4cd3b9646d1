mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do python bench.py --steps 300 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6)"; done
