#!/usr/bin/env python
"""Isolated timing of the non-GEMM kernels at network shapes (L2 flushed between launches).
   python tools/op_bench.py [dwgram|attn|gate|finish|all] [reps]
Also the command to hand to ncu (-k regex:dwgram ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P3 = lib.PREC_BF16X3


def rnd(*shape):
    return torch.randn(*shape, device=dev, dtype=torch.float32)


def timeit(name, fn, bytes_, inner=10):
    """GPU time per launch: `inner` back-to-back launches replayed from a CUDA graph (no CPU launch cost in the
    interval); the operands of the big shapes exceed L2, the tiny kernels run L2-warm as they do in the network."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    ts = []
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {1e3*t:8.1f} us  {bytes_/t/1e6:7.0f} GB/s")


# (B, H, W, C, heads) of the cube512 workload: L1, L2, L3, decoder/refinement
SHAPES = [(1, 512, 512, 64, 2), (1, 256, 256, 128, 4), (1, 128, 128, 256, 8), (1, 512, 512, 128, 2)]
SMALL = [(16, 64, 64, 64, 2), (16, 32, 32, 128, 4), (16, 16, 16, 256, 8), (16, 64, 64, 128, 2)]
for (B, H, W, C, heads) in SHAPES + SMALL:
    N = B * H * W
    c = C // heads
    if which in ("dwgram", "all") and lib.dwgram_supported(C, c):
        t3, w9, v = rnd(N, 3 * C), rnd(9, 3 * C), rnd(N, C)
        nfl, nch = lib.dwgram_partial_floats(B, heads, c, H, W)
        part = torch.empty(nfl, device=dev)
        timeit(f"dwgram B{B} {H}x{W} C{C} h{heads}", lambda: lib.dwgram(View.of(t3), w9, View.of(v), part, B, H, W, C, heads, P3), 16.0 * N * C)
        if which in ("finish", "all"):
            scratch = torch.empty(B * heads * (c * c + 2 * c), device=dev)
            temp, wout = torch.ones(heads, device=dev), rnd(C, C)
            img = torch.empty(B, lib.bimg_bytes(C, C), dtype=torch.uint8, device=dev)
            timeit(f"finish B{B} C{C} h{heads} chunks{nch}", lambda: lib.spectral_finish(part, nch, scratch, temp, wout, None, img, B, heads, c), 4.0 * nfl)
        del t3, v
    if which in ("attn", "all"):
        qkv, out, wm = rnd(N, 3 * C), rnd(N, C), torch.empty(N // 64 * C, device=dev)
        bias = rnd(heads, 64, 64)
        for shift in (0, 4):
            timeit(f"attn B{B} {H}x{W} C{C} h{heads} s{shift}", lambda: lib.window_attn(View.of(qkv), bias, View.of(out), wm, B, H, W, C, heads, shift, P3), 16.0 * N * C)
        del qkv, out
    if which in ("gate", "all"):
        r = 8 if C != 128 or heads != 2 else 16
        B_ = N // 64
        w = {"promptT": rnd(C, 128), "promptb": rnd(128), "downT": rnd(C, r), "downb": rnd(r), "param": rnd(128, r),
             "qT": rnd(r, r), "kvT": rnd(r, 2 * r), "p2T": rnd(r, r), "p2b": rnd(r), "upT": rnd(r, C)}
        cm, gate = rnd(B_, C), torch.empty(B_, C, device=dev)
        timeit(f"gate B_{B_} C{C} r{r}", lambda: lib.local_gate(cm, w, gate, B_, C, r), 8.0 * B_ * C)
        logits = rnd(B_, 144)
        timeit(f"gate_tail B_{B_} C{C} r{r}", lambda: lib.local_gate_tail(View.of(logits), w, gate, B_, C, r), 4.0 * B_ * (C + 144))
        wcat = lib.Weight(None, lib.pack_bimg(rnd(128 + r, C), 128 + r, C), 128 + r, C)
        bias = rnd(128 + r)
        timeit(f"gate_gemm B_{B_} C{C} N{128 + r}", lambda: lib.gemm(View.of(cm), wcat, View.of(logits), 128 + r, bias=bias, precision=P3), 4.0 * B_ * (C + 144))
