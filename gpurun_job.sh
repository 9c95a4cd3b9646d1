mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for L in 32 8; do
echo "c3 lanes=$L $(PAINTRL_MOVE_LANES=$L python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['roofline']['frac'])")"
done 2>&1 | tee gpurun_out/sweep_c3.txt
