"""Micro-benchmark of the two weight-gradient kernels on the shapes of one training step (CUDA events, L2 flushed).
   python tools/wgrad_bench.py [bf16|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View

prec = lib.PREC_BF16X3 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else lib.PREC_BF16
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib.load()
N1, N2, N3 = 32 * 64 * 64, 32 * 32 * 32, 32 * 16 * 16
SHAPES = [  # (name, M, O, I, kwargs)
    ("L1 qkv 192x64", N1, 192, 64, {}), ("L1 fc1 352x64", N1, 352, 64, {}), ("L1 fc2 64x176", N1, 64, 176, {}),
    ("L1 proj 64x64", N1, 64, 64, {}),
    ("D1 qkv 384x128", N1, 384, 128, {}), ("D1 fc1 704x128", N1, 704, 128, {}), ("D1 fc2 128x352", N1, 128, 352, {}),
    ("L2 fc1 704x128", N2, 704, 128, {}), ("L3 fc1 1376x256", N3, 1376, 256, {}), ("L3 qkv 768x256", N3, 768, 256, {}),
    ("conv out 32x128 taps", N1, 32, 128, dict(taps=9, H=64, W=64, so=9 * 128, si=9, st=1)),
    ("conv up2_1 256x128 taps", N2, 256, 128, dict(taps=9, H=32, W=32, so=9 * 128, si=9, st=1)),
    ("spectral P 128x128 per-sample", N1, 128, 128, dict(rows_per_batch=4096, dw_batch_stride=128 * 128, so=128)),
    ("gate 128x64 windows", N1 // 64, 128, 64, {}),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"{'shape':34s} {'tcgen05 us':>11s} {'mma.sync us':>12s} {'GB/s tc':>9s} {'GB/s mma':>9s}")
for name, M, O, I, kw in SHAPES:
    dY = torch.randn(M, O, device=dev)
    X = torch.randn(M, I, device=dev)
    nout = (M // kw["rows_per_batch"]) * O * I if "rows_per_batch" in kw else O * I * (9 if kw.get("taps") else 1)
    dW = torch.zeros(nout, device=dev)
    res = []
    for eng in (1, 0):
        lib.load().mphsir_debug_wgrad_tc(eng)
        for _ in range(2):
            lib.wgrad(View.of(dY), View.of(X), dW, prec, **kw)
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.wgrad(View.of(dY), View.of(X), dW, prec, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts)[len(ts) // 2])
    by = 4.0 * M * (O + I) * (1 if not kw.get("taps") else 1)
    print(f"{name:34s} {res[0]:11.1f} {res[1]:12.1f} {by / res[0] / 1e3:9.0f} {by / res[1] / 1e3:9.0f}")
lib.load().mphsir_debug_wgrad_tc(-1)
