mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
PAINTRL_DEBUG=1 timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/debug.txt | tee gpurun_out/bench_c2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['roofline']['ray_full_scans_per_env_step'])"
grep "move cells" gpurun_out/debug.txt | head -2
PAINTRL_TRACE=1 python -m paintrl_b200.build --force >/dev/null 2>&1
timeout 300 python profiles/timeline.py --out gpurun_out/trace_c2.npy 2>&1 | grep -v "move duration, [0-9]" | tee gpurun_out/timeline_c2.txt
PAINTRL_PROFILE=1 python -m paintrl_b200.build --force >/dev/null 2>&1
python profiles/phase_profile.py --envs 4096 --steps 30 --dump-rays gpurun_out/slow_rays.npy 2>&1 | tail -1
