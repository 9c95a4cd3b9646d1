set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 300 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c2.json
python bench.py --steps 300 --warmup 10 --no-flush --no-cpu-baseline > gpurun_out/bench_c2_warm.json 2>> gpurun_out/bench_c2.err
python bench.py --steps 100 --warmup 5 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/prof_r01_v2 python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_v2.log 2>&1
ls -la gpurun_out
