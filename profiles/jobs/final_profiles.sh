# Round-end evidence for profiles/: counters of every workload, full ncu captures of the dominant kernels (summarised
# on the box: the reports are too large to travel), timelines.  Run AFTER the last change to csrc/ (the counters carry
# the source hash).
mkdir -p gpurun_out
python profiles/capture_counters.py --tag r02 --workloads c2,c3,c3_late,c4,c5,c2_normal 2>&1 | tail -8
B="python bench.py --no-extra --no-cpu-baseline --no-parity --e2e-sync"
cap() { tag=$1; envs=$2; shift 2; ncu --set full --clock-control none --import-source on -o /tmp/$tag -f "$@" > gpurun_out/ncu_$tag.log 2>&1; bash profiles/summarize_rep.sh /tmp/$tag.ncu-rep $tag $envs; rm -f /tmp/$tag.ncu-rep; }
cap r02_c2_step_fused 4096 -k regex:step_fused -s 12 -c 1 $B --workload c2 --steps 12 --warmup 5
cap r02_c3_move_paint 16384 -k regex:'move_fast|paint_kernel' -s 12 -c 2 $B --workload c3_late --steps 8 --warmup 4
cap r02_c4_step_fused 8192 -k regex:step_fused -s 6 -c 1 $B --workload c4 --steps 5 --warmup 3
cap r02_c5_move_paint 65536 -k regex:'move_fast|paint_kernel' -s 10 -c 2 $B --workload c5 --steps 6 --warmup 3
cap r02_policy_act 65536 -k regex:policy_act -s 3 -c 1 python bench.py --workload c5 --rollout --steps 1 --warmup 1 --no-graph --no-extra --no-cpu-baseline --no-parity
cap r02_normal_paint 4096 -k regex:paint_normal -s 4 -c 1 $B --workload c2_normal --steps 4 --warmup 3
PAINTRL_TRACE=1 python -m paintrl_b200.build --force 2>&1 | tail -1
PAINTRL_FUSED=1 python profiles/timeline.py --steps 40 > gpurun_out/r02_c2_timeline.txt 2>&1; tail -3 gpurun_out/r02_c2_timeline.txt | cut -c1-200
python profiles/timeline.py --workload c3_late --envs 16384 --steps 20 > gpurun_out/r02_c3_timeline.txt 2>&1
du -sh gpurun_out; ls gpurun_out | head -50
