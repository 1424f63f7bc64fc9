#!/usr/bin/env python
"""Top stalled SASS instructions of one launch in an .ncu-rep (needs --import-source on / -lineinfo).
   python tools/ncu_hot.py gpurun_out/prof.ncu-rep <launch-index> [topN]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[idx]:starts[idx + 1]]
print(rows[0][1][:100])
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall_cols}
print("samples", tot, "| stall mix:", ", ".join(f"{c[6:]}={100*v/tot:.0f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for r in sorted(data, key=lambda r: -int(r[S] or 0))[:topn]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(r[ix["Address"]][-5:], "%6d %5.1f%%" % (int(r[S]), 100 * int(r[S]) / tot), r[ix["Source"]][:80], st)
