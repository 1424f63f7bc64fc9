#!/usr/bin/env python
"""Micro-benchmark of the local spectral gate kernels: warm (weights L2-resident) vs cold (256 MB flush before every call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View
dev = "cuda"
for B_, C, r in ((4096, 128, 16), (1024, 128, 16), (256, 256, 32), (4096, 64, 8)):
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g).to(dev)
    w = {"promptT": rn(C, 128) * C ** -0.5, "promptb": rn(128), "downT": rn(C, r) * C ** -0.5, "downb": rn(r),
         "param": rn(128, r), "qT": rn(r, r), "kvT": rn(r, 2 * r), "p2T": rn(r, r), "p2b": rn(r), "upT": rn(r, C)}
    cm = rn(B_, C)
    gate = torch.empty(B_, C, device=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    ldl = (128 + r + 15) // 16 * 16
    logits = rn(B_, ldl)
    for name, fn in (("gate2", lambda: lib.local_gate2(cm, w, gate, B_, C, r)),
                     ("tail", lambda: lib.local_gate_tail(View.of(logits), w, gate, B_, C, r)),
                     ("gate(old)", lambda: lib.local_gate(cm, w, gate, B_, C, r))):
        for cold in (False, True):
            ts = []
            for _ in range(12):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts = sorted(ts[2:])
            print(f"B_={B_:5d} C={C:3d} r={r:2d} {name:10s} {'cold' if cold else 'warm'}: median {ts[len(ts)//2]:6.1f} us  min {ts[0]:6.1f} us")

