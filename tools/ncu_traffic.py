#!/usr/bin/env python
"""Per-kernel-family DRAM traffic of one profiled step from an ncu CSV log captured with
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
   python tools/ncu_traffic.py gpurun_out/traffic.csv > profiles/rNN_traffic.json
bench.py reads the JSON to fill roofline.traffic (bytes per launch of the dominant kernel family)."""
import csv, io, json, re, sys
from collections import defaultdict

text = open(sys.argv[1]).read()
rd = csv.DictReader(io.StringIO(text[text.find('"ID"'):]))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
per = defaultdict(lambda: {"launches": set(), "dram_bytes": 0.0, "us": 0.0})
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mphsir::", "").replace("void ", "")
    fam = re.sub(r"<.*", "", name).split("::")[-1]
    if fam in ("gemm_tc_kernel",):
        fam = "gemm_tc_kernel"
    v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r.get("Metric Unit", ""), 1.0)
    e = per[fam]
    e["launches"].add(r["ID"])
    if r["Metric Name"].startswith("dram__bytes"):
        e["dram_bytes"] += v
    elif r["Metric Name"] == "gpu__time_duration.sum":
        e["us"] += v
out = {k: {"launches": len(v["launches"]), "dram_bytes_per_launch": v["dram_bytes"] / len(v["launches"]),
           "dram_bytes_total": v["dram_bytes"], "kernel_us_total": v["us"]} for k, v in per.items()}
json.dump({"source": sys.argv[1], "note": "ncu --clock-control none, one step, dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "families": dict(sorted(out.items(), key=lambda kv: -kv[1]["kernel_us_total"]))}, sys.stdout, indent=1)
