mkdir -p gpurun_out
PAINTRL_TRACE=1 python -m paintrl_b200.build --force 2>&1 | tail -1
PAINTRL_FUSED=1 python profiles/timeline.py --steps 40 2>&1 | grep -E "move duration|their|slowest|/" | cut -c1-600
