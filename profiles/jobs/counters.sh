# ncu counters of every bench workload for the current sources + one full capture of the C2 kernel
mkdir -p gpurun_out
python profiles/capture_counters.py --tag r02 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:step_fused -s 12 -c 1 -o gpurun_out/r02_final_c2 -f python bench.py --steps 12 --warmup 5 --no-extra --no-cpu-baseline --no-parity --e2e-sync > gpurun_out/ncu_r02_final_c2.log 2>&1
ls -la gpurun_out/r02_final_c2.ncu-rep gpurun_out/kernel_counters.json
