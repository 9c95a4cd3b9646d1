mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"paint_kernel" -s 6 -c 1 -o gpurun_out/prof_r01v5_c3 -f python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
