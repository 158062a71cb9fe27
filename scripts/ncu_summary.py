#!/usr/bin/env python
"""summarise an .ncu-rep (raw page) into the handful of metrics quoted in DESIGN.md / profiles/"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_bytes_pipe_lsu_mem_global', 'lts__t_bytes.sum ', 'sm__inst_executed_pipe_fp64', 'smsp__inst_executed_op_shared']
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys):
            if 'stalled' in h and not h.endswith('per_issue_active.ratio'):
                continue
            print("  %-95s %-14s %s" % (h, u, v))
