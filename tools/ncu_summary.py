#!/usr/bin/env python
"""Condenses an .ncu-rep (one kernel launch, --set full) into the few lines profiles/*.md quote.
Usage: python tools/ncu_summary.py gpurun_out/agg.ncu-rep [rows_for_per_row_stats]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(*args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("--page", "raw")
hdr, units, vals = raw[0], raw[1], raw[2]
get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for k in keys:
    if k in get:
        print("%-75s %s %s" % (k, get[k][0], get[k][1]))
src = page("--page", "source", "--print-source", "sass")
hi = [i for i, r in enumerate(src) if "Instructions Executed" in r][0]
h = src[hi]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
c = Counter()
tot = 0
for r in src[hi + 1:]:
    if len(r) > ia and r[ia].isdigit():
        s = r[isrc].strip()
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        c[op] += int(r[ia])
        tot += int(r[ia])
print("warp instructions executed: %d" % tot + (" (%.1f thread-instructions per row)" % (tot * 32 / rows_n) if rows_n else ""))
print("top opcodes: " + ", ".join("%s %.1f%%" % (op, 100.0 * n / tot) for op, n in c.most_common(12)))
