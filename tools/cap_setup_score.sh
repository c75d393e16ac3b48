O=gpurun_out
python tools/one_rollout.py > $O/r2c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_score|k_setup_synth" -c 2 -f -o $O/r2c_setup_score python tools/one_rollout.py > $O/r2c_ncu.log 2>&1
ncu -i $O/r2c_setup_score.ncu-rep --page raw --csv > $O/r2c_raw.csv 2>/dev/null
python tools/ncu_rows.py $O/r2c_raw.csv > $O/r2c_setup_score_summary.txt
ncu -i $O/r2c_setup_score.ncu-rep --page source --csv --print-source cuda,sass > $O/r2c_src.csv 2>/dev/null
python tools/ncu_lines.py $O/r2c_src.csv 60 > $O/r2c_setup_score_lines.txt; rm -f $O/r2c_src.csv $O/r2c_raw.csv
