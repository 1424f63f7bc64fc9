#!/usr/bin/env python
"""Micro-benchmark + role cycle counters of the fused MLP kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib, engine as E
from mp_hsir_b200.lib import View, Weight
dev = "cuda"
M = 262144
FLAGS = [int(a) for a in sys.argv[1:] if a.isdigit()] or [0]
PRECS = [(lib.PREC_BF16, "x1")] if "x1" in sys.argv else [(lib.PREC_BF16X3, "x3")]
for C, hid in (((64, 170),) if "c64" in sys.argv else ((128, 340),)):
    hp = E._ceil(hid, 16)
    x = torch.randn(M, C, device=dev)
    w1 = torch.randn(2 * hp, C, device=dev) * C ** -0.5
    w2 = torch.randn(C, hp, device=dev) * hp ** -0.5
    W1 = Weight(None, lib.pack_bimg(w1, 2 * hp, C), 2 * hp, C)
    W2 = Weight(None, lib.pack_bimg(w2, C, hp), C, hp)
    b1, b2 = torch.randn(2 * hp, device=dev), torch.randn(C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y = torch.empty(M, C, device=dev)
    for prec, pn, fl in [(pr, pn_, f) for pr, pn_ in PRECS for f in FLAGS]:
        lib.load().mphsir_debug_mlp_flags(fl)
        run = lambda: lib.mlp(View.of(x), (g, b), W1, b1, W2, b2, View.of(y), hp, prec)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
        lib.load().mphsir_debug_mlp_counters(dbg.data_ptr()); run(); torch.cuda.synchronize(); lib.load().mphsir_debug_mlp_counters(None)
        d = dbg.double().mean(0).tolist()
        names = ["MMA.total", "-", "MMA.w_x", "MMA.w_b", "MMA.w_h", "MMA.w_acc2", "GLU.total", "GLU.w_acc1", "EPI.total", "EPI.w_acc2", "EPI.tmem_ld", "CVT.total", "CVT.w_x_empty", "GLU.ld", "GLU.math", "GLU.st"]
        print(f"flags={fl} C={C} hid={hid} {pn}: {ms*1e3:.1f} us  {2.0*M*C*3*hp/ms/1e9:.1f} TF  {12.0*M*C/ms/1e6:.0f} GB/s")
        print("     " + "  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, d)))
