B="python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --no-parity --e2e-sync"
run() { name=$1; shift; env "$@" $B $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s value %.1fM us/step %.1f e2e %.1fM' % ('$name', d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['step_us'])
"; }
for P in 1 0 1 0; do
EXTRA="--workload c2"; run pf${P}_flush PAINTRL_L2_PREFETCH=$P
done
EXTRA="--workload c2 --no-flush"; run pf1_noflush PAINTRL_L2_PREFETCH=1
EXTRA="--workload c2 --no-flush"; run pf0_noflush PAINTRL_L2_PREFETCH=0
python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle_batch.py -m gpu -x -q 2>&1 | tail -3
