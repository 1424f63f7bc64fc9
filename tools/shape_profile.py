#!/usr/bin/env python
"""Per-shape kernel time breakdown of one forward (CUDA events around every launch)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mp_hsir_b200 import lib

workload = sys.argv[1] if len(sys.argv) > 1 else "cube512"
model, shape, unit, units, _, _ = bench.WORKLOADS[workload]
dev = torch.device("cuda", 0)
if "--no-psplit" in sys.argv:
    sys.argv.remove("--no-psplit")
    lib.load().mphsir_debug_tc_psplit(0)
if "--no-pair" in sys.argv:
    sys.argv.remove("--no-pair")
    lib.load().mphsir_debug_tc_cluster(0)
cfg, net = bench.build_net(model, dev)
if "--no-split-gate" in sys.argv:
    sys.argv.remove("--no-split-gate")
    net.engine().split_gate = False
if len(sys.argv) > 2:
    net.set_precision(sys.argv[2])
x, _, tid = bench.make_input(shape, 0, workload)
x, tid = x.to(dev), tid.to(dev)

# wrap lib.gemm / conv3x3 cost lambdas to carry shapes in the tag
orig_launch = lib._launch
def launch(what, call, cost=None):
    if lib.PROFILER is not None and cost is not None and what in ("gemm_fwd", "conv3x3_fwd", "window_attn_fwd", "dwconv3x3_fwd", "gram_partial_fwd"):
        fl, by, tag = cost()
        cost = (lambda fl=fl, by=by, tag=tag: (fl, by, f"{tag} fl={fl/1e9:.2f}G by={by/1e6:.0f}MB"))
    return orig_launch(what, call, cost)
lib._launch = launch
with torch.no_grad():
    for _ in range(3):
        net(x, tid)
    lib.PROFILER = lib.Profiler()
    net(x, tid)
    agg = lib.PROFILER.summary()
    lib.PROFILER = None
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot:.2f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:70]:
    print(f"{a['ms']:7.3f} ms x{a['launches']:3d}  {a['bytes']/a['ms']/1e6:7.0f} GB/s {a['flops']/a['ms']/1e9:7.1f} TF  {k}")
