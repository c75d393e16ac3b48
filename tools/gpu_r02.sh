#!/bin/bash
# Round-2 capture script (runs on the GPU box under gpurun): every ncu run follows a plain run of the same command.
set -u
O=gpurun_out
mkdir -p $O
RAW='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput|sm__warps_active|launch__registers_per_thread|smsp__inst_executed.sum|smsp__issue_active|smsp__sass_average_branch_targets_threads_uniform|smsp__thread_inst_executed_per_inst_executed|sm__inst_executed_pipe_alu|sm__inst_executed_pipe_fma|lts__t_sector_hit_rate|sm__throughput|launch__grid_size|launch__occupancy_limit'
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $O/$name "$@" > $O/${name}_ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  python tools/ncu_rows.py $O/${name}_raw.csv "$RAW" > $O/${name}_summary.txt 2>&1
}
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-large --no-sub"
$B > $O/r2_launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches.csv $B > $O/r2_launch_ncu.log 2>&1
cap r2_step_1m   k_step 8 4  python tools/one_rollout.py
cap r2_step_8m   k_step 8 4  python tools/one_rollout.py 8388608
cap r2_step_m17  k_step 8 4  python tools/one_rollout.py 1048576 17
cap r2_step_m17b k_step 24 4 python tools/one_rollout.py 1048576 17
cap r2_setup_score "k_score|k_setup_synth" 0 2 python tools/one_rollout.py
cap r2_step_forced k_step 56 4 python tools/one_rollout_forced.py
ls -la $O/*.ncu-rep
# keep the reports small enough to travel: only the 1M step report keeps its source view
rm -f $O/r2_step_8m.ncu-rep $O/r2_step_m17.ncu-rep $O/r2_step_m17b.ncu-rep $O/r2_step_forced.ncu-rep
