mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py | tee gpurun_out/bench_r01_final_c2.json | cut -c1-300
python bench.py --workload c3 --steps 200 --no-cpu-baseline | tee gpurun_out/bench_r01_final_c3.json | cut -c1-120
python bench.py --workload c4 --steps 50 --no-cpu-baseline | tee gpurun_out/bench_r01_final_c4.json | cut -c1-120
python bench.py --workload c5 --steps 100 --no-cpu-baseline | tee gpurun_out/bench_r01_final_c5.json | cut -c1-120
python bench.py --workload c5 --rollout --steps 3 | tee gpurun_out/bench_r01_final_c5_rollout.json | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_r01_final_c2.csv python bench.py --steps 60 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
