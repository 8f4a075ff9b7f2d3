#!/usr/bin/env python
"""Selected raw-page rows of an .ncu-rep, one block per profiled launch: python tools/ncu_rows.py report.ncu-rep [extra metric ...]
(the text summaries under profiles/ are made with this; needs `ncu` on PATH, no GPU)."""
import csv
import io
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"] + sys.argv[2:]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(head)}
for r in rows[2:]:
    print("Kernel Name".ljust(76), r[col["Kernel Name"]][:110])
    for w in WANT:
        if w in col:
            print(w.ljust(76), units[col[w]].ljust(16), r[col[w]])
    print()
