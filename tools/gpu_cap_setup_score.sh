set -u
O=gpurun_out
python tools/one_rollout.py > $O/r19_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_score|k_setup_synth" -s 0 -c 2 -f -o $O/r19_setup_score python tools/one_rollout.py > $O/r19_ncu.log 2>&1
ncu -i $O/r19_setup_score.ncu-rep --page raw --csv > $O/r19_raw.csv 2>/dev/null
python tools/ncu_rows.py $O/r19_raw.csv > $O/r19_setup_score_summary.txt 2>&1
ncu -i $O/r19_setup_score.ncu-rep --page source --csv --print-source cuda,sass > $O/r19_src.csv 2>/dev/null
python tools/ncu_lines.py $O/r19_src.csv 40 > $O/r19_setup_score_lines.txt 2>&1
rm -f $O/r19_src.csv $O/r19_raw.csv
cat $O/r19_setup_score_summary.txt
