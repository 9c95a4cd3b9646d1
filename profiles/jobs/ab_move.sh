mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
B="python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; env "$@" $B $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('%s value %.1fM us/step %.1f e2e %.1fM bail %.4f parity %s launches %d' % ('$name', d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, r['move_bailouts_per_env_step'], d.get('parity_check',{}).get('ok'), d['gpu_launches']))
"; }
EXTRA="--workload c2"
run c2_two_kernel PAINTRL_FUSED=0
run c2_fused PAINTRL_FUSED=1
EXTRA="--workload c2 --envs 8192"
run c2_8192_two_kernel PAINTRL_FUSED=0
run c2_8192_fused PAINTRL_FUSED=1
EXTRA="--workload c2 --envs 16384"
run c2_16384_two_kernel PAINTRL_FUSED=0
run c2_16384_fused PAINTRL_FUSED=1
EXTRA="--workload c5 --steps 60"
run c5_two_kernel PAINTRL_FUSED=0
run c5_fused PAINTRL_FUSED=1
EXTRA="--workload c3_late --steps 60"
run c3l_two_kernel PAINTRL_FUSED=0
run c3l_fused PAINTRL_FUSED=1
EXTRA="--workload c3 --steps 60"
run c3_two_kernel PAINTRL_FUSED=0
run c3_fused PAINTRL_FUSED=1
EXTRA="--workload c4 --steps 30"
run c4_two_kernel PAINTRL_FUSED=0
run c4_fused PAINTRL_FUSED=1
