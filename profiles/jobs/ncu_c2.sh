mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 5 --no-extra --no-cpu-baseline --no-parity --e2e-sync"
PAINTRL_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:step_fused -s 12 -c 1 -o gpurun_out/r02_fused_c2 -f $B > gpurun_out/ncu_fused.log 2>&1
PAINTRL_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:'move_fast|paint_kernel' -s 24 -c 2 -o gpurun_out/r02_two_c2 -f $B > gpurun_out/ncu_two.log 2>&1
tail -3 gpurun_out/ncu_fused.log gpurun_out/ncu_two.log
ls -la gpurun_out/*.ncu-rep
