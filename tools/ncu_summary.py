#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
  python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep          > profiles/rNN_kernel.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    return name.replace("mphsir::", "").replace("void ", "")


def launches(path):
    rows = []
    with open(path) as f:
        text = f.read()
    start = text.find('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), v * scale))
    # keep exactly one forward step: the launches between two consecutive nchw_to_tokens kernels
    marks = [i for i, (k, _) in enumerate(rows) if "nchw_to_tokens" in k]
    if len(marks) >= 2 and "all" not in sys.argv[3:]:
        rows = rows[marks[0]:marks[1]]
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms summed kernel time (ncu: cold-cache, serialised -> compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {us:12.1f} {us / tot:7.3f} {us / n:10.1f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print(f"## {short(r[idx['Kernel Name']])}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx:
                print(f"   {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
