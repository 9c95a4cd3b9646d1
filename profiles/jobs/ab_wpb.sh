B="python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --no-parity --e2e-sync"
run() { name=$1; shift; env "$@" $B $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s value %.1fM us/step %.1f' % ('$name', d['value']/1e6, d['ms_per_step']*1e3), d['step_us'])
"; }
for W in 2 4; do
PAINTRL_NVCC_EXTRA="-DPAINTRL_FUSED_WPB=$W" python -m paintrl_b200.build --force 2>&1 | tail -1
EXTRA="--workload c2"; run wpb$W A=1
EXTRA="--workload c2"; run wpb${W}_again A=1
EXTRA="--workload c4 --steps 20"; run wpb${W}_c4 A=1
done
