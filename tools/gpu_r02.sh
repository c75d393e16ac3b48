#!/bin/bash
# Round-2 capture script (runs on the GPU box under gpurun): every ncu run follows a plain run of the same command.
set -u
O=gpurun_out
mkdir -p $O
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $O/$name "$@" > $O/${name}_ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  python tools/ncu_rows.py $O/${name}_raw.csv > $O/${name}_summary.txt 2>&1
  ncu -i $O/$name.ncu-rep --page source --csv --print-source cuda,sass > $O/${name}_src.csv 2>/dev/null
  python tools/ncu_lines.py $O/${name}_src.csv 40 > $O/${name}_lines.txt 2>&1
  rm -f $O/${name}_src.csv
}
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-large --no-sub"
$B > $O/r2_launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches.csv $B > $O/r2_launch_ncu.log 2>&1
cap r2b_step_1m   k_step 8 4  python tools/one_rollout.py
cap r2b_step_8m   k_step 8 4  python tools/one_rollout.py 8388608
cap r2b_step_m17  k_step 24 4 python tools/one_rollout.py 1048576 17
cap r2b_setup_score "k_score|k_setup_synth" 0 2 python tools/one_rollout.py
cap r2b_step_forced k_step 56 4 python tools/one_rollout_forced.py
cap r2b_obs "k_obs_expand_all|k_select_action_all|k_bucket" 40 5 python tools/one_selfplay.py
ls -la $O/*.ncu-rep
rm -f $O/r2b_step_8m.ncu-rep $O/r2b_step_m17.ncu-rep $O/r2b_step_forced.ncu-rep $O/r2b_obs.ncu-rep
