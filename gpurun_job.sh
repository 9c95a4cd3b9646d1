# The standard GPU-box job of this repo:  gpurun --timeout 1500 -- 'bash gpurun_job.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 300 --warmup 10 | tee gpurun_out/bench_c2.json | cut -c1-400
