B="python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --e2e-sync"
run() { name=$1; shift; env "$@" $B $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s value %.1fM us/step %.1f parity %s bail %.5f' % ('$name', d['value']/1e6, d['ms_per_step']*1e3, d.get('parity_check',{}).get('ok'), d['roofline']['move_bailouts_per_env_step']), d['step_us'])
"; }
EXTRA="--workload c2"; run c2 A=1
EXTRA="--workload c2"; run c2_again A=1
EXTRA="--workload c5 --steps 60"; run c5 A=1
