set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/prof_r01_v2 python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_v2.log 2>&1
