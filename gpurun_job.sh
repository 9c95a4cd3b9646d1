# The standard GPU-box job of this repo:  gpurun --timeout 1800 -- 'bash gpurun_job.sh'
mkdir -p gpurun_out
nvidia-smi -L | head -2
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
( time python bench.py --steps 300 --warmup 10 ) 2>&1 | tee gpurun_out/bench_default.json | cut -c1-1800
python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --no-parity --e2e-sync | tee gpurun_out/bench_e2e_sync.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sync e2e', d['e2e']['value']/1e6, 'device', d['value']/1e6)"
