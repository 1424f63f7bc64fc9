#!/usr/bin/env python
"""Role cycle counters of the tcgen05 window-attention kernel (mp_hsir_b200/csrc/window_attn_tc.cu).
   python tools/attn_bench.py [C heads H W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View
a = [int(v) for v in sys.argv[1:5]] or [128, 2, 512, 512]
C, heads, H, W = a
dev = "cuda"
qkv = torch.randn(H * W, 3 * C, device=dev)
bias = torch.randn(heads, 64, 64, device=dev)
out = torch.empty(H * W, C, device=dev)
wm = torch.empty(H * W // 64, C, device=dev)
bias_t = bias.transpose(1, 2).contiguous()
run = lambda: lib.window_attn(View.of(qkv), bias, View.of(out), wm, 1, H, W, C, heads, 4, precision=lib.PREC_BF16X3, bias_t=bias_t)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"C={C} heads={heads} {H}x{W}: {e0.elapsed_time(e1)*100:.1f} us")
cnt = torch.zeros(148, 16, dtype=torch.int64, device=dev)
lib.load().mphsir_debug_window_attn_tc_counters(cnt.data_ptr()); run(); torch.cuda.synchronize()
lib.load().mphsir_debug_window_attn_tc_counters(None)
d = cnt.double().mean(0).tolist()
names = ["TMA.total", "TMA.wait_empty", "MMA.total", "MMA.issue", "S0.total", "S0.wait_s", "S0.wait_o", "S0.softmax", "S0.epilogue",
         "S1.total", "S1.wait_s", "S1.wait_o", "S1.softmax", "S1.epilogue", "CV.wait_landing", "CV.wait_image"]
print("  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, d)))
