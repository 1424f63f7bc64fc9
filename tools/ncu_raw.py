#!/usr/bin/env python
"""Selected raw metrics per launch of an .ncu-rep:  python tools/ncu_raw.py rep [substr ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"] + extra
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("----", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            print(f"  {h:88s} {r[i]:>16s} {units[i]}")
