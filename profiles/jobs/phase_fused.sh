mkdir -p gpurun_out
PAINTRL_PROFILE=1 python -m paintrl_b200.build --force 2>&1 | tail -2
echo "=== fused, L2 flushed"; PAINTRL_FUSED=1 python profiles/phase_profile.py --steps 40 2>&1 | tail -32 | tee gpurun_out/phase_fused_flush.txt
echo "=== fused, warm L2"; PAINTRL_FUSED=1 python profiles/phase_profile.py --steps 40 --no-flush 2>&1 | tail -32 | tee gpurun_out/phase_fused_warm.txt
