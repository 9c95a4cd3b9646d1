bash profiles/jobs/loader.sh
B="python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --no-parity --e2e-sync"
run() { name=$1; shift; env "$@" $B $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s value %.1fM us/step %.1f' % ('$name', d['value']/1e6, d['ms_per_step']*1e3), d['step_us'])
"; }
EXTRA="--workload c2 --flush-mode write"; run write A=1
EXTRA="--workload c2 --flush-mode write+read"; run write_read A=1
EXTRA="--workload c2 --flush-mode write"; run write A=1
EXTRA="--workload c2 --flush-mode write+read"; run write_read A=1
EXTRA="--workload c5 --steps 60 --flush-mode write"; run c5_write A=1
EXTRA="--workload c5 --steps 60 --flush-mode write+read"; run c5_write_read A=1
