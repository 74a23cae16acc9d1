#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the handful of counters DESIGN.md and bench.py quote: duration, DRAM bytes, issue
utilisation, SIMT efficiency, occupancy limiters, stall mix.  Usage: ncu_summary.py <report.ncu-rep> [> profiles/xyz.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr, units = rows[0], rows[1]
M = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_static", "static smem/block"),
     ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
     ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
     ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
     ("sm__inst_executed.avg.per_cycle_active", "IPC per SM"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
     ("lts__t_sector_hit_rate.pct", "L2 hit %")]
ST = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "no_instruction", "branch_resolving", "mio_throttle", "math_pipe_throttle",
      "not_selected", "lg_throttle", "dispatch_stall"]
print("# %s  (ncu --set full --clock-control none; per launch)" % rep.split("/")[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("\n## %s" % d.get("Kernel Name", "?").split("(")[0])
    for k, name in M:
        if k in d and d[k] != "":
            print("  %-34s %s %s" % (name, d[k], units[hdr.index(k)]))
    st = []
    for s in ST:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if k in d and d[k] not in ("", "0"):
            st.append("%s %.2f" % (s, float(d[k])))
    print("  stall cycles per issued instruction:  " + ", ".join(st))
