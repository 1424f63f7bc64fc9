#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM engine on the network's shapes (CUDA events, L2 flushed by size)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View, Weight

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 262144
shapes = [  # name, K, N, epi, ln, extra
    ("proj   K128 N128 bias", 128, 128, lib.EPI_BIAS, False),
    ("sqkv   K128 N384", 128, 384, lib.EPI_BIAS, False),
    ("qkv    K128 N384 ln", 128, 384, lib.EPI_BIAS, True),
    ("fc1    K128 N704 ln glu", 128, 704, lib.EPI_GLU, True),
    ("fc2    K352 N128 res", 352, 128, lib.EPI_RESIDUAL, False),
    ("projf  K128 N512 proj", 128, 512, lib.EPI_PROJ, False),
    ("apply  K128 N128 res", 128, 128, lib.EPI_RESIDUAL, False),
    ("projf64 K64 N256 proj", 64, 256, lib.EPI_PROJ, False),
    ("proj64 K64 N64", 64, 64, lib.EPI_BIAS, False),
    ("qkv64  K64 N192 ln", 64, 192, lib.EPI_BIAS, True),
]
if "--no-pair" in sys.argv:
    lib.load().mphsir_debug_tc_cluster(0)
if "--no-ebox1" in sys.argv:
    lib.load().mphsir_debug_tc_ebox1(0)
for name, K, N, epi, ln in shapes:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * K ** -0.5
    img = lib.pack_bimg(w, N, K)
    W = Weight(None, img, N, K)
    n_out = N // 2 if epi == lib.EPI_GLU else N
    y = torch.empty(M, n_out, device=dev)
    bias = torch.randn(N, device=dev)
    g, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    res = torch.randn(M, N, device=dev) if epi == lib.EPI_RESIDUAL else None
    proj = epi == lib.EPI_PROJ
    if proj:
        Cc = N // 4
        res = torch.randn(M, Cc, device=dev)
        y = torch.empty(M, Cc, device=dev)
        y2 = torch.empty(M, 3 * Cc, device=dev)
        gate = torch.randn(M // 64, Cc, device=dev)
        side = int(M ** 0.5)
    for prec, pn in ((lib.PREC_BF16X3, "x3"), (lib.PREC_BF16, "x1")):
        def run():
            if proj:
                lib.gemm(View.of(a), W, View.of(y), N, bias=bias, epi=epi, res1=View.of(res), gate=gate, Y2=View.of(y2),
                         n_split=Cc, H=side, W=side, shift=4, rows_per_batch=M, precision=prec)
                return
            lib.gemm(View.of(a), W, View.of(y), N, bias=bias, ln=(g, b) if ln else None, epi=epi,
                     res1=View.of(res) if res is not None else None, precision=prec)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if "--profile" in sys.argv:
            dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
            lib.load().mphsir_debug_tc_counters(dbg.data_ptr())
            run()
            torch.cuda.synchronize()
            lib.load().mphsir_debug_tc_counters(None)
            d = dbg.double().mean(0).tolist()
            names = ["Bload.total", "Bload.wait_empty", "MMA.total", "MMA.wait_acc", "MMA.wait_A", "MMA.wait_B", "EPI.total",
                     "EPI.wait_full", "EPI.tmem_ld", "CV.total", "CV.bar", "CV.wait_tma", "MMA.issue", "MMA.commit", "-"]
            print("      " + "  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, d)))
        byts = 4.0 * (M * K + M * n_out + (res.numel() if res is not None else 0))
        print(f"{name:28s} {pn}  {ms*1e3:8.1f} us  {byts/ms/1e6:7.1f} GB/s  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
