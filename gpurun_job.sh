set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
PAINTRL_DEBUG=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --steps 100 --warmup 5 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
