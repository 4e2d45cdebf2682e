#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per profiled launch with the metrics
B200_PROFILING.md names.  usage: ncu_summary.py raw.csv [more-metric-substrings...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
extra = sys.argv[2:]
cols = [i for i, h in enumerate(hdr) if h in want or any(e in h for e in extra)]
stall = [i for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    print('-' * 100)
    for i in cols:
        print("  %-78s %s %s" % (hdr[i], r[i][:90], units[i]))
    st = sorted(((float(r[i].replace(',', '') or 0), hdr[i]) for i in stall), reverse=True)
    tot = sum(v for v, _ in st) or 1.0
    for v, h in st[:6]:
        print("  stall(pc samples) %-60s %5.1f %%" % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot))
