set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --steps 100 --warmup 5 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c3.json
