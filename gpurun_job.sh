mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | tee gpurun_out/bench_c2.json | cut -c1-200
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/bench_c4.json | cut -c1-1500
