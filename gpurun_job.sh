set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
PAINTRL_MOVE_LANES=32 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu32.log
for L in 8 16 32; do
PAINTRL_DEBUG=1 PAINTRL_MOVE_LANES=$L python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2_$L.json 2> gpurun_out/bench_c2_$L.err; tail -1 gpurun_out/bench_c2_$L.json | cut -c1-120; tail -2 gpurun_out/bench_c2_$L.err
done
python bench.py --steps 100 --warmup 5 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c3.json | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
