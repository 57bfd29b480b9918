#!/usr/bin/env python
"""Text summary of an ncu report for committing under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
usage: python profiles/summarize.py <report.ncu-rep> > profiles/rNN_<kernel>.txt"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
print("# ncu summary of %s (captured with --set full --clock-control none; times under ncu are not benchmark values)" % os.path.basename(rep))
for w in WANT:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %-12s %s" % (h, units[i], vals[i]))
print()
print("# hottest source lines (profiles/ncu_lines.py)")
sys.stdout.flush()
subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), rep, "16"])
