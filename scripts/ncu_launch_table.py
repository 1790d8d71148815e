#!/usr/bin/env python
"""Markdown table from an ncu launch list (`ncu --metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv --log-file X`):
per kernel the launch count, the average duration over the last `tail` launches and its share of the sum of those averages.
With `total` as the fourth argument the shares are of the summed durations / instructions of ALL launches (a run whose kernels launch a
different number of times each, e.g. a proving call).
usage: ncu_launch_table.py launches.csv [tail=6] [skip-regex] [total]"""
import csv
import re
import sys

path = sys.argv[1]
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 6
skip = re.compile(sys.argv[3]) if len(sys.argv) > 3 else re.compile(r"k_mb_|k_from_uniform|k_gens|k_fb_bases|k_fb_fill|at::|vectorized|elementwise|k_l2")
rows = [ln for ln in open(path) if ln.startswith('"')]
dur, inst = {}, {}
for r in csv.DictReader(rows):
    name = r["Kernel Name"].split("(")[0]
    if skip.search(name):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    (dur if r["Metric Name"] == "gpu__time_duration.sum" else inst).setdefault(name, []).append(v)
total_mode = len(sys.argv) > 4 and sys.argv[4] == "total"
if total_mode:
    print("| kernel | launches | avg µs | total µs | share of kernel time | warp instr (M, all launches) | share of instr |")
    print("|---|---|---|---|---|---|---|")
    tt, ti = sum(sum(v) for v in dur.values()), sum(sum(v) for v in inst.values()) or 1.0
    for k in sorted(dur, key=lambda k: -sum(dur[k])):
        print("| `%s` | %d | %.1f | %.0f | %.1f %% | %.2f | %.1f %% |" % (k, len(dur[k]), sum(dur[k]) / len(dur[k]) / 1e3, sum(dur[k]) / 1e3, 100 * sum(dur[k]) / tt,
                                                                  sum(inst.get(k, [0])) / 1e6, 100 * sum(inst.get(k, [0])) / ti))
    print("\nkernel time of all launches: %.0f µs; %.1f M warp instructions" % (tt / 1e3, ti / 1e6))
    sys.exit(0)
avg = {k: sum(v[-tail:]) / len(v[-tail:]) for k, v in dur.items()}
tot = sum(avg.values())
ins = {k: sum(v[-tail:]) / len(v[-tail:]) for k, v in inst.items()}
itot = sum(ins.values()) or 1.0
print("| kernel | launches | avg µs | share of step time | warp instr (M) | share of instr |")
print("|---|---|---|---|---|---|")
for k in sorted(avg, key=lambda k: -avg[k]):
    print("| `%s` | %d | %.1f | %.1f %% | %.2f | %.1f %% |" % (k, len(dur[k]), avg[k] / 1e3, 100 * avg[k] / tot, ins.get(k, 0) / 1e6, 100 * ins.get(k, 0) / itot))
print("\nsum of average launch times: %.0f µs; %.1f M warp instructions" % (tot / 1e3, itot / 1e6 if ins else 0))
