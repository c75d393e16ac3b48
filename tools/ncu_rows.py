"""Condenses `ncu --page raw --csv` into one line per metric with one column per profiled launch.
   python tools/ncu_rows.py raw.csv [regex]   (default: the metrics DESIGN.md / bench.py quote)"""
import csv, re, sys
DEFAULT = ("^(dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
           "sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
           "launch__occupancy_limit_registers|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|"
           "smsp__sass_average_branch_targets_threads_uniform.pct|smsp__thread_inst_executed_per_inst_executed.ratio|"
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active|"
           "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active|lts__t_sector_hit_rate.pct|"
           "sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__inst_executed.avg.per_cycle_elapsed|sm__cycles_elapsed.avg|"
           "smsp__cycles_active.avg|l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum|l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum)$")
path = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else DEFAULT)
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
data = rows[hdr + 2:]
kn = names.index("Kernel Name")
print("kernels:", [re.sub(r"\(.*", "", r[kn]).replace("void tk::", "").replace("void ", "") for r in data])
for j, m in enumerate(names):
    if pat.search(m):
        print("%-66s %-12s %s" % (m, units[j], "  ".join(r[j] for r in data)))
