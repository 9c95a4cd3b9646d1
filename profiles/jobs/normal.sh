timeout 900 python -m pytest tests/test_gpu_normal_paint.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c2_normal --steps 20 --warmup 3 --no-extra --no-cpu-baseline --e2e-sync 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2_normal value %.2fM ms/step %.3f parity %s' % (d['value']/1e6, d['ms_per_step'], d.get('parity_check',{}).get('ok')))"
python bench.py --workload c2_normal --envs 16384 --steps 10 --warmup 3 --no-extra --no-cpu-baseline --e2e-sync --no-parity 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2_normal 16384 envs value %.2fM ms/step %.3f' % (d['value']/1e6, d['ms_per_step']))"
