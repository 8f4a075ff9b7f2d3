#!/usr/bin/env python
"""Digest of an ncu report (read in the build container): headline counters, dynamic instruction mix per update and
the top stall sites.   python tools/ncu_digest.py <file.ncu-rep> <updates per launch> [top]"""
import collections
import csv
import subprocess
import sys

rep, updates = sys.argv[1], float(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
for h, v in zip(hdr, vals):
    if any(h == w or h.startswith(w) for w in want) and "per_second" not in h and "pct_of_peak" not in h.replace("dram__throughput", ""):
        print("%-70s %s" % (h, v))
    if "issue_stalled" in h and "per_issue_active" in h and float(v or 0) > 0.05:
        print("%-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall/"), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iS, iE, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot, stot, lines = collections.Counter(), collections.Counter(), 0, 0, []
for r in rows[2:]:
    try:
        e, s = int(r[iE]), int(r[iP])
    except (ValueError, IndexError):
        continue
    t = r[iS].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += e
    samp[op] += s
    tot += e
    stot += s
    lines.append((s, e, r[iS].strip()))
print("warp instructions per update: %.1f" % (tot / updates))
for op, e in ops.most_common(top_n):
    print("  %-12s %6.2f /update  %5.1f%% of samples" % (op, e / updates, 100.0 * samp[op] / max(stot, 1)))
print("top stall sites:")
for s, e, txt in sorted(lines, reverse=True)[:top_n]:
    print("  %5.1f%%  x%.2f/upd  %s" % (100.0 * s / max(stot, 1), e / updates, txt[:90]))
