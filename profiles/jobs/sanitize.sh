mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/r02_sanitizer_racecheck.txt
grep -E "Race reported|Warning|Error:" gpurun_out/r02_sanitizer_racecheck.txt | sed -E 's/0x[0-9a-f]+/ADDR/g' | sort | uniq -c | sort -rn | head -20
grep -A6 -m3 "Race reported\|Warning: " gpurun_out/r02_sanitizer_racecheck.txt | head -60
