mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle_batch.py tests/test_gpu_raster.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do python bench.py --workload c3 --steps 200 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6)"; done
