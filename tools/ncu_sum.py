#!/usr/bin/env python
"""Short per-kernel summary of an .ncu-rep (`ncu -i X --page raw --csv` piped through)."""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            v = r[hdr.index(w)]
            if w.startswith("smsp__average_warps"):
                if float(v or 0) < 0.15:
                    continue
                w = w.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "")
            print(w, v[:80])
