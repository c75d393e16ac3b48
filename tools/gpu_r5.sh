set -u
O=gpurun_out
python tools/one_rollout.py > $O/r5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_step" -s 8 -c 4 -f -o $O/r5_step python tools/one_rollout.py > $O/r5_ncu.log 2>&1
ncu -i $O/r5_step.ncu-rep --page raw --csv > $O/r5_step_raw.csv 2>/dev/null
python tools/ncu_rows.py $O/r5_step_raw.csv > $O/r5_step_summary.txt 2>&1
ncu -i $O/r5_step.ncu-rep --page source --csv --print-source cuda,sass > $O/r5_step_src.csv 2>/dev/null
python tools/ncu_lines.py $O/r5_step_src.csv 400 > $O/r5_step_lines.txt 2>&1
gzip -f $O/r5_step_src.csv
rm -f $O/r5_step_raw.csv
cat $O/r5_step_summary.txt
