#!/usr/bin/env python
"""Summarise an ncu --set full report (read here, no GPU): tools/ncu_summary.py rep.ncu-rep [title]"""
import csv, subprocess, sys, io
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'sm__cycles_elapsed.avg.per_second']
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"## {title}\n\n| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} | {d[k][1]} |")
    print()
