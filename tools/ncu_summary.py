#!/usr/bin/env python3
"""Print the headline metrics of an ncu report (one kernel launch). usage: ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr, units, vals = rows[0], rows[1], rows[-1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'launch__waves_per_multiprocessor',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum']
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print("%-70s %-14s %s" % (w, units[i], vals[i]))
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
        try:
            if float(vals[i]) > 4: print("%-70s %-14s %s" % (h, units[i], vals[i]))
        except ValueError: pass
