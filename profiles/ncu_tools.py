"""Helpers to summarise ncu exports (ncu -i X.ncu-rep --page raw|source --csv) into profiles/*.txt."""
import collections
import csv
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_registers',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum')


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for i, h in enumerate(hdr):
        if h in KEEP:
            print('%-72s %-16s %s' % (h, units[i], ' '.join(r[i] for r in rows[2:])))
    print('kernels:', [r[ki][:50] for r in rows[2:]])


def source(path, top=20):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    isrc, iex, ist = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == 'Kernel Name':
            break
        try:
            data.append((r[isrc], int(r[iex]), int(r[ist])))
        except ValueError:
            pass
    print(len(data), 'SASS instructions; executed (warp-level)', sum(d[1] for d in data), '; stall samples', sum(d[2] for d in data))
    op, st = collections.Counter(), collections.Counter()
    for s, e, t in data:
        m = s.split()
        name = (m[1] if m and m[0].startswith('@') else (m[0] if m else '')).split('.')[0]
        op[name] += e
        st[name] += t
    tot = sum(op.values())
    for k, v in op.most_common(14):
        print('  %-10s executed %9d (%.3f)  stall samples %d' % (k, v, v / tot, st[k]))
    print('top stall sites:')
    for s, e, t in sorted(data, key=lambda d: -d[2])[:top]:
        print('  %5d  %s' % (t, s[:110]))


if __name__ == '__main__':
    {'raw': raw, 'source': source}[sys.argv[1]](sys.argv[2])
