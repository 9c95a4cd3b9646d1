timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c5 --rollout --steps 3 --warmup 1 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 rollout value %.1fM us/step %.1f' % (d['value']/1e6, d['ms_per_step']*1e3))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:policy_act -s 5 -c 3 python bench.py --workload c5 --rollout --steps 1 --warmup 1 --no-graph --no-extra --no-cpu-baseline --no-parity 2>&1 | grep -E "gpu__time_duration|policy_act" | head -6
