# summarize_rep.sh REP TAG ENVS: key metrics, per-line tables and SASS opcode mix of every kernel in an .ncu-rep -> gpurun_out/TAG_*.txt
REP=$1; TAG=$2; ENVS=$3; D=$(dirname $0)
ncu -i $REP --page raw --csv > /tmp/$TAG.raw.csv 2>/dev/null
python $D/ncu_raw.py /tmp/$TAG.raw.csv > gpurun_out/${TAG}_key_metrics.txt
ncu -i $REP --page source --csv --print-source cuda,sass > /tmp/$TAG.src.csv 2>/dev/null
python $D/ncu_lines.py /tmp/$TAG.src.csv 40 > gpurun_out/${TAG}_lines.txt 2>&1
ncu -i $REP --page source --csv --print-source sass > /tmp/$TAG.sass.csv 2>/dev/null
python $D/ncu_sass_ops.py /tmp/$TAG.sass.csv $ENVS > gpurun_out/${TAG}_sass_ops.txt 2>&1
