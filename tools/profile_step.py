"""One profiled step between cudaProfilerStart/Stop (run under `ncu --profile-from-start off ...`).
   python tools/profile_step.py infer|train|metrics [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_net, make_input, make_train_batch

mode = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else ("fp32" if mode == "infer" else "bf16")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if mode == "infer":
    cfg, net = build_net("natural", dev)
    net.set_precision(prec)
    x = make_input((1, 31, 512, 512), 0, "cube512")[0].to(dev)
    tid = torch.zeros(1, dtype=torch.long, device=dev)
    with torch.no_grad():
        for _ in range(2):
            net(x, tid)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        net(x, tid)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
elif mode == "metrics":
    # evaluation metrics of one restored cube + degradation of one training batch (mp_hsir_b200/csrc/metrics.cu)
    from mp_hsir_b200 import metrics as GM
    from mp_hsir_b200.degrade import degrade_batch
    noisy, clean, _ = make_input((1, 31, 512, 512), 0, "cube512")
    noisy, clean = noisy.to(dev), clean.to(dev)
    batch = torch.rand(32, 31, 64, 64, device=dev)
    for _ in range(2):
        GM.compute_psnr_ssim(noisy, clean)
        degrade_batch(batch, seed=1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    GM.compute_psnr_ssim(noisy, clean)
    degrade_batch(batch, seed=2)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    cfg, net = build_net("natural", dev)
    net.set_precision(prec)
    net.train()
    with torch.no_grad():
        net.output.weight.mul_(0.05)
    tr = net.trainer()
    noisy, clean, tid = (t.to(dev) for t in make_train_batch((32, 31, 64, 64), 0, cfg.task_classes))
    for _ in range(2):
        tr.train_step(noisy, clean, tid)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.train_step(noisy, clean, tid)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
