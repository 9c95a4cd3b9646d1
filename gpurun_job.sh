mkdir -p gpurun_out
python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-flush | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 warm L2', d['value']/1e6, d['ms_per_step']*1e3)"
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 flushed', d['value']/1e6, d['ms_per_step']*1e3)"
PAINTRL_TRACE=1 python -m paintrl_b200.build --force >/dev/null 2>&1
timeout 300 python profiles/timeline.py --workload c3 --envs 16384 2>&1 | grep -v "move duration, " | tee gpurun_out/timeline_c3.txt
