mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
(time python bench.py) 2>&1 | tee gpurun_out/bench_default.json | cut -c1-2500
(time python bench.py --impl reference --steps 5 --warmup 1) 2>&1 | tee gpurun_out/bench_reference.json | cut -c1-900
