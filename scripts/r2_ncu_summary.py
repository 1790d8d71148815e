#!/usr/bin/env python
"""profiles/r02_ncu_summary.md from the two ncu captures of scripts/r2_profile.sh (launch list CSV + `--set full` report).
usage: python scripts/r2_ncu_summary.py gpurun_out/r2_launches.csv gpurun_out/r2_prof_full.ncu-rep|_raw.csv [title] > profiles/r02_ncu_summary.md
(the report is read through `ncu -i ... --page raw --csv`; a CSV made by that command on the GPU box is taken as it is)"""
import csv
import io
import subprocess
import sys

launches, rep = sys.argv[1], sys.argv[2]
table = subprocess.run([sys.executable, "scripts/ncu_launch_table.py", launches, "6"], capture_output=True, text=True).stdout
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
title = sys.argv[3] if len(sys.argv) > 3 else "16-job pass, one pass in flight"
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
M = [("gpu__time_duration.sum", "duration µs", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__block_size", "block", 1.0),
     ("launch__registers_per_thread", "regs / thread", 1.0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %", 1.0),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %", 1.0),
     ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA / IMAD pipe cycles active %", 1.0),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", 1.0),
     ("smsp__warps_active.avg.per_cycle_active", "warps resident per scheduler", 1.0), ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible per scheduler per cycle", 1.0),
     ("lts__t_sector_hit_rate.pct", "L2 hit rate %", 1.0),
     ("smsp__inst_executed.sum", "warp instructions (M)", 1e-6), ("dram__bytes_read.sum", "DRAM read (MB)", None), ("dram__bytes_write.sum", "DRAM write (MB)", None),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait", 1.0),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle", 1.0),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard", 1.0),
     ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch", 1.0),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected", 1.0),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier", 1.0)]
units = dict(zip(hdr, rows[1]))
kern = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    n = d["Kernel Name"].split("(")[0].replace("void ", "")
    kern.setdefault(n, d)            # first captured launch of each kernel


def val(d, key, scale):
    try:
        x = float(d[key].replace(",", ""))
    except (KeyError, ValueError):
        return "-"
    if scale is None:                 # bytes with a unit column
        u = units.get(key, "")
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        return "%.2f" % x
    x *= scale
    return "%.3g" % x if abs(x) < 1000 else "%.0f" % x


names = sorted(kern, key=lambda n: -float(kern[n]["gpu__time_duration.sum"].replace(",", "")))
print("## 2. `--set full` captures (first captured launch of each kernel; %s)\n" % title)
print("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
print("|---|" + "---|" * len(names))
for key, label, scale in M:
    if key in hdr:
        print("| %s | " % label + " | ".join(val(kern[n], key, scale) for n in names) + " |")
print("\n## 1. Launch list (`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum`, last 6 launches of each kernel)\n")
print(table)
