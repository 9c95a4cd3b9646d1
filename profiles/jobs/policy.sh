timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -25
B="python bench.py --workload c5 --rollout --steps 3 --warmup 3"
$B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('native rollout', d['value']/1e6, 'M, us/step', d['ms_per_step']*1e3, d['config']['workload'][-220:])"
