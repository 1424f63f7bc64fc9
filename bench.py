#!/usr/bin/env python
"""bench.py — headline benchmark of the MP-HSIR hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cube512|patch16|rs256|train64]

A *step* is one forward pass of the drop-in ``MP_HSIR_Net`` over one batch of synthetic input
(BASELINE.json: "HSI cubes/s (31x512x512 infer)").  Default workload ``cube512`` = natural-scene
model, one 1x31x512x512 cube per step per GPU, task 0 (Gaussian denoise); N>1 shards independent
cubes over ranks (weak scaling, no data-path collective).  Prints ONE JSON line (rank 0).

  value     : whole-job cubes/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : same metric through the public host-to-host API (mp_hsir_b200.pipeline.HostPipeline.restore_stream)
              with pinned HOST buffers: H2D of every cube and D2H of every restored cube inside the timed region,
              overlapped with the forward of the neighbouring cube on copy streams
  roofline  : dominant kernel family from an instrumented pass (CUDA events around every launch,
              outside the timed region), algorithmic FLOPs/bytes from mp_hsir_b200.lib cost model
  cpu_baseline : the oracle port (oracle/mp_hsir_oracle.py, PyTorch CPU, all host threads) on a
              bounded sample of the same workload

``--impl reference`` times the reference's CPU implementation of the path on the host cores at the SAME shape: the
unmodified net/MP_HSIR.py from the git-ignored oracle/_ref (oracle/build_ref.py, travels to the GPU box); the oracle
port only when that copy is missing.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (model, per-GPU input shape, unit name, units per step per GPU, cpu sample shape, sample fraction)
    "cube512": ("natural", (1, 31, 512, 512), "cubes/s", 1, (1, 31, 256, 256), 0.25),
    "patch16": ("natural", (16, 31, 64, 64), "patches/s", 16, (4, 31, 64, 64), 4.0),
    "rs256": ("remote_sensing", (1, 100, 256, 256), "patches/s", 1, (1, 100, 128, 128), 0.25),
    # BASELINE config 4: one optimisation step (forward, clamp+L1, backward, AdamW; NCCL gradient all-reduce for N>1)
    "train64": ("natural", (32, 31, 64, 64), "patches/s", 32, (2, 31, 64, 64), 2.0),
}
# committed outputs of the UNMODIFIED reference at exactly these shapes (oracle/make_golden.py --big): rank 0 runs the
# fixture's input, and its timed output is checked against the fixture before a number is printed
GOLDEN_CASE = {"cube512": "nat_cube512", "patch16": "nat_b16_64", "rs256": "rs_b1_256"}
PARITY_TOL = {"fp32": 1e-4, "fp32_exact": 1e-4, "bf16": 1e-2}   # north_star: max|d|/max|ref|
METRIC = {"cube512": "HSI cubes/s (31x512x512 infer)", "patch16": "HSI patches/s (16x31x64x64 infer)",
          "rs256": "RS patches/s (100x256x256 infer)", "train64": "train patches/s (31x64x64, batch 32/GPU)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(workload: str, family: str):
    """DRAM bytes per launch of a kernel family from the committed ncu capture of this workload (profiles/traffic_<workload>.json,
    written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one step)."""
    p = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    try:
        with open(p) as f:
            fam = json.load(f)["families"].get(family)
        return None if fam is None else {"bytes_per_launch": fam["dram_bytes_per_launch"], "launches": fam["launches"],
                                         "source": os.path.relpath(p, ROOT)}
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_net(model: str, device):
    from mp_hsir_b200 import MP_HSIR_Net
    from mp_hsir_b200.config import NetConfig
    from mp_hsir_b200.synth import fill_state_dict_
    cfg = NetConfig.natural() if model == "natural" else NetConfig.remote_sensing()
    net = MP_HSIR_Net(cfg.in_channel, cfg.out_channel, cfg.dim, task_classes=cfg.task_classes)
    fill_state_dict_(net, seed=0)   # random-init weights of the named architecture (no checkpoints offline)
    return cfg, net.to(device).eval()


def golden_meta(workload: str):
    with open(os.path.join(ROOT, "tests", "golden", "cases.json")) as f:
        return json.load(f)[GOLDEN_CASE[workload]]


def make_input(shape, seed, workload=None):
    """(input, clean or None, task ids).  With a workload name: the recipe of its golden fixture (seed 0 = the fixture's
    own input; other ranks draw other seeds of the same recipe)."""
    from mp_hsir_b200.synth import synthetic_input, synthetic_scene
    if workload in GOLDEN_CASE:
        meta = golden_meta(workload)
        tid = torch.tensor(meta["task_id"], dtype=torch.long)
        if meta["recipe"][0] == "scene":
            noisy, clean = synthetic_scene(meta["recipe"][1], meta["recipe"][2], seed=seed)
            return noisy.contiguous(), clean, tid
        return synthetic_input(tuple(meta["recipe"][1]), seed=seed), None, tid
    if shape[2] >= 256 and shape[0] == 1:
        noisy, clean = synthetic_scene(shape[1], shape[2], seed=seed)
        return noisy.contiguous(), clean, torch.zeros(shape[0], dtype=torch.long)
    return synthetic_input(shape, seed=seed), None, torch.zeros(shape[0], dtype=torch.long)


def psnr_per_band(y, clean) -> float:
    """utils/val_utils.py:49-69 of the reference: per-band PSNR (data_range 1) on clip(.,0,1), mean over bands."""
    y, c = y.clamp(0, 1).double(), clean.clamp(0, 1).double()
    return float((10.0 * torch.log10(1.0 / ((y - c) ** 2).mean(dim=(-1, -2)))).mean())


def golden_check(workload: str, y: torch.Tensor, clean, precision: str):
    """The timed output against the committed reference output of this exact input (strided subsample of every band +
    all-pixel band sums + PSNR against the clean cube).  Raises SystemExit on a miss: a fast wrong answer is not a result."""
    import numpy as np
    meta = golden_meta(workload)
    g = np.load(os.path.join(ROOT, "tests", "golden", GOLDEN_CASE[workload] + ".npz"))
    s = meta["stride"]
    yc = y.detach().cpu().double()
    e_sub = float((yc[:, :, ::s, ::s] - torch.from_numpy(g["sub"]).double()).abs().max() / meta["out_absmax"])
    hw = y.shape[-1] * y.shape[-2]
    e_mean = float((yc.sum(dim=(-1, -2)) - torch.from_numpy(g["band_sums"]).double()).abs().max() / hw / meta["out_absmax"])
    res = {"fixture": f"tests/golden/{GOLDEN_CASE[workload]}.npz (unmodified reference, oracle/make_golden.py --big)",
           "max_abs_err_over_max_abs_ref": e_sub, "band_mean_err_over_max_abs_ref": e_mean, "tolerance": PARITY_TOL[precision]}
    if clean is not None and "psnr_ref_vs_clean" in meta:
        res["psnr_delta_db"] = abs(psnr_per_band(yc, clean) - meta["psnr_ref_vs_clean"])
    ok = e_sub < PARITY_TOL[precision] and e_mean < PARITY_TOL[precision] and res.get("psnr_delta_db", 0.0) <= 0.01
    if not ok or not bool(torch.isfinite(y).all()):
        raise SystemExit(f"bench: output does not match the reference fixture: {res}")
    return res


def reference_forward_fn(model: str):
    """-> (callable(x, tid) running the reference's CPU implementation of the path, kind).  kind "reference": the
    UNMODIFIED net/MP_HSIR.py from oracle/_ref (oracle/build_ref.py; travels to the GPU box); "port": the oracle
    restatement, when that copy is missing."""
    from mp_hsir_b200.config import NetConfig
    cfg = NetConfig.natural() if model == "natural" else NetConfig.remote_sensing()
    from oracle import ref_import
    if ref_import.available():
        net = ref_import.build_reference(cfg, seed=0)
        return (lambda x, tid: net(x, tid)), "reference"
    from mp_hsir_b200 import MP_HSIR_Net
    from mp_hsir_b200.synth import synth_tensor, synthetic_clip_prompt
    from oracle import mp_hsir_oracle as O
    shapes = MP_HSIR_Net(cfg.in_channel, cfg.out_channel, cfg.dim, task_classes=cfg.task_classes)
    sd = {k: synth_tensor(k, p.shape, 0) for k, p in shapes.named_parameters()}
    clip = synthetic_clip_prompt(cfg.task_classes)
    return (lambda x, tid: O.forward(sd, cfg, x, tid, clip)), "port"


def cpu_baseline(model: str, sample_shape, frac: float, unit: str, steps: int = 1, warmup: int = 0, workload=None,
                 budget_s: float = 1e9):
    """The reference's CPU implementation of the path timed on the host cores (reported, not the target).  With
    `workload`: the FULL workload shape (one step = one whole batch, frac = units per step); else a bounded sample.
    Stops early once `budget_s` seconds of timed work are spent (at least one timed step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind = reference_forward_fn(model)
    x, _, tid = make_input(sample_shape, 0, workload)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd(x, tid)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            if i >= warmup and sum(times) > budget_s:
                break
    t = sum(times) / len(times)
    what = "the unmodified reference module (oracle/_ref/net/MP_HSIR.py, eval, no_grad)" if kind == "reference" else "the oracle port"
    return {"value": frac / t, "unit": unit, "cores": cores, "kind": kind, "steps_timed": len(times),
            "sample": f"{len(times)} forward(s) of {what} on {list(x.shape)} fp32 "
                      f"(= {frac:g} step-units each), {t:.2f} s per forward, torch CPU {cores} threads"}, t


# ------------------------------------------------------------------------------------------------
# training workload (BASELINE config 4)
# ------------------------------------------------------------------------------------------------


def make_train_batch(shape, seed, T):
    """clean patches + ONE random degradation per sample, drawn like ImageTransformDataset.__getitem__ of the reference
    (utils/dataset_utils.py:128-146) from the array-only recipes of its de_dict (:112): Gaussian noise sigma ~ U(30,70)/255,
    non-iid Gaussian noise with a per-band sigma from {10,30,50,70}/255, random pixel mask (ratio 0.7/0.8/0.9,
    degradation_utils.py:227-233) and band loss (ratio 0.1/0.2/0.3).  The task id [B,1] is the degradation id, as in the
    training collate (utils/dataset_utils.py:140)."""
    from mp_hsir_b200.synth import synthetic_input
    g = torch.Generator().manual_seed(1000 + seed)
    B, C, H, W = shape
    clean = synthetic_input(shape, seed=seed)
    noisy = clean.clone()
    tid = torch.randint(0, min(T, 4), (B, 1), generator=g)
    for b in range(B):
        k = int(tid[b, 0])
        if k == 0:      # gaussianN
            sigma = (30.0 + 40.0 * torch.rand(1, generator=g)) / 255.0
            noisy[b] += sigma * torch.randn(C, H, W, generator=g)
        elif k == 1:    # non-iid noise (the array-only part of complexN)
            sig = torch.tensor([10.0, 30.0, 50.0, 70.0])[torch.randint(0, 4, (C,), generator=g)] / 255.0
            noisy[b] += sig.view(C, 1, 1) * torch.randn(C, H, W, generator=g)
        elif k == 2:    # inpaint: random mask
            ratio = (0.7, 0.8, 0.9)[int(torch.randint(0, 3, (1,), generator=g))]
            noisy[b] *= (torch.rand(C, H, W, generator=g) > ratio).float()
        else:           # bandmiss
            ratio = (0.1, 0.2, 0.3)[int(torch.randint(0, 3, (1,), generator=g))]
            lost = torch.randperm(C, generator=g)[: max(1, int(round(ratio * C)))]
            noisy[b, lost] = 0.0
    return noisy.contiguous(), clean.contiguous(), tid


def cpu_train_baseline(model: str, sample_shape, steps: int = 1, warmup: int = 0):
    """The reference's CPU training step — forward (train mode) + clamp/L1 (train.py:58-61) + autograd backward +
    torch.optim.AdamW (train.py:69) — on the host cores, bounded sample.  Unmodified reference module from oracle/_ref when
    present (kind "reference"), else the oracle port's autograd (kind "port")."""
    from mp_hsir_b200.config import NetConfig
    from mp_hsir_b200.synth import synth_tensor, synthetic_clip_prompt
    from mp_hsir_b200 import MP_HSIR_Net
    from oracle import ref_import
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = NetConfig.natural() if model == "natural" else NetConfig.remote_sensing()
    noisy, clean, tid = make_train_batch(sample_shape, 0, cfg.task_classes)
    if ref_import.available():
        kind = "reference"
        net = ref_import.build_reference(cfg, seed=0).train()
        params = list(net.parameters())
        fwd = lambda: net(noisy, tid)  # noqa: E731
    else:
        kind = "port"
        from oracle import mp_hsir_oracle as O
        shapes = MP_HSIR_Net(cfg.in_channel, cfg.out_channel, cfg.dim, task_classes=cfg.task_classes)
        sd = {k: synth_tensor(k, p.shape, 0).requires_grad_(True) for k, p in shapes.named_parameters()
              if "text_linear" not in k and "clip_linear" not in k}
        clip = synthetic_clip_prompt(cfg.task_classes)
        params = list(sd.values())
        fwd = lambda: O.forward(sd, cfg, noisy, tid, clip)  # noqa: E731
    opt = torch.optim.AdamW(params, lr=2e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(fwd().clamp(0, 1), clean)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    what = "the unmodified reference module (oracle/_ref, train mode)" if kind == "reference" else "the oracle port"
    return {"value": sample_shape[0] / t, "unit": "patches/s", "cores": cores, "kind": kind,
            "sample": f"{len(times)} training step(s) of {what} (autograd + torch AdamW) on a batch of "
                      f"{sample_shape[0]} {list(sample_shape[1:])} patches fp32, {t:.2f} s per step, torch CPU {cores} threads"}, t


def run_train(args, embedded: bool = False):
    """train64 workload.  embedded: called from the default (cube512) run, which owns the process group and prints the
    line itself; then `args.workload` is still cube512 and the training precision is the training default (bf16)."""
    import torch.distributed as dist
    from mp_hsir_b200 import lib
    from mp_hsir_b200.parallel import max_over_ranks as _max

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not embedded:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    workload = "train64"
    model, shape, unit, units, sample_shape, _ = WORKLOADS[workload]
    precision = "bf16" if embedded else args.precision
    cfg, net = build_net(model, device)
    net.set_precision(precision)
    net.train()
    with torch.no_grad():
        net.output.weight.mul_(0.05)  # random-init output conv scaled so restored values sit in the clamp's live range
    tr = net.trainer(lr=2e-4)
    from mp_hsir_b200.parallel import all_reduce_gradients
    all_reduce = all_reduce_gradients if world > 1 else None
    # every rank trains on its own shard of the global batch (DDP batch sharding, train.py:118)
    noisy_h, clean_h, tid_h = (t.pin_memory() for t in make_train_batch(shape, rank, cfg.task_classes))
    noisy_d, clean_d, tid_d = noisy_h.to(device), clean_h.to(device), tid_h.to(device)
    loss_h = torch.zeros(1).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = args.cuda_graph != "off"

    def step(x, c, t):
        return tr.train_step(x, c, t, world_size=world, all_reduce=all_reduce, cuda_graph=use_graph)

    losses = []
    for _ in range(args.warmup):
        losses.append(float(step(noisy_d, clean_d, tid_d)))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.LAUNCHES
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(noisy_d, clean_d, tid_d)
    e1.record()
    barrier()
    launches = lib.LAUNCHES - n0
    ms_dev = _max(e0.elapsed_time(e1), device)
    losses.append(float(loss))
    if not all(l == l and l < 1e3 for l in losses):
        raise SystemExit(f"bench: training diverged, losses {losses}")
    # ---- e2e: pinned host batch -> device, loss back to the host, every step ---------------------
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2e_losses = []

    # host-fed step: only the CLEAN batch crosses PCIe; the degradation of every sample is synthesised on the device
    # (mp_hsir_b200.degrade: host draws the per-sample parameters, one mphsir_degrade launch draws the noise / masks)
    from mp_hsir_b200 import degrade as DG
    dgen = torch.Generator().manual_seed(4321 + rank)
    h2d_extra = [0]

    def e2e_step():
        tid_s, sigma_s, keep_s, ratio_s = DG.draw_parameters(shape[0], shape[1], DG.RECIPES, dgen)
        clean_s = clean_h.to(device, non_blocking=True)
        noisy_s = DG.degrade(clean_s, sigma_s, keep_s, ratio_s, seed=len(e2e_losses) + 1000 * rank)
        h2d_extra[0] = (sigma_s.numel() + keep_s.numel() + ratio_s.numel()) * 4 + tid_s.numel() * 8
        loss = step(noisy_s, clean_s, tid_s.to(device, non_blocking=True))
        loss_h.copy_(loss, non_blocking=True)
        e2e_losses.append(loss.clone())

    e2e_step()  # untimed: first-use allocations of the host-fed path (cudaMalloc of the staging tensors) are not a steady-state cost
    barrier()
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = _max(f0.elapsed_time(f1), device)
    losses.append(float(loss_h))
    losses_e2e = [float(l) for l in e2e_losses]
    clocks = sampler.stop() if sampler else None
    graph_ms = None
    if use_graph:
        tr.time_graphs = True
        step(noisy_d, clean_d, tid_d)
        tr.time_graphs = False
        graph_ms = {k: round(v, 3) for k, v in (tr.graph_ms or {}).items()}

    roof, breakdown = None, None
    agg = None
    if not args.no_roofline:
        # the instrumented step contains the gradient all-reduce: EVERY rank takes it (rank 0 with per-launch events)
        if rank == 0:
            lib.PROFILER = lib.Profiler()
        step(noisy_d, clean_d, tid_d)
        if rank == 0:
            agg = lib.PROFILER.summary()
            lib.PROFILER = None
        barrier()
    if agg is not None:
        pk = peaks()
        total_ms = sum(a["ms"] for a in agg.values())
        breakdown = {k: {"ms_per_step": round(a["ms"], 3), "share": round(a["ms"] / total_ms, 4), "launches_per_step": a["launches"],
                         "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2) if a["ms"] else 0,
                         "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] else 0}
                     for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

        def family(tag):
            if tag.startswith(("gemm_tc", "conv3x3_tc")):
                return "gemm_tc_kernel"
            return "wgrad_kernel" if tag.startswith("wgrad") else tag
        fam = {}
        for k, a in agg.items():
            f = fam.setdefault(family(k), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            for key in f:
                f[key] += a[key]
        name, a = max(fam.items(), key=lambda kv: kv[1]["ms"])
        sec = a["ms"] * 1e-3
        mma = 3.0 if (precision == "fp32" and name in ("gemm_tc_kernel", "wgrad_kernel")) else 1.0
        t_hbm = a["bytes"] / (pk["hbm_gbs"] * 1e9)
        t_tensor = a["flops"] * mma / (pk["tf_sustained"] * 1e12)
        if t_tensor > t_hbm:
            ach = a["flops"] * mma / sec / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": None}
        else:
            ach = a["bytes"] / sec / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None}
        tr_ = ncu_traffic(workload, name) if precision == "bf16" else None
        if tr_ is not None:
            roof["traffic"] = tr_["bytes_per_launch"]
            roof["traffic_source"] = f"{tr_['source']}: ncu dram read+write bytes / launch over the {tr_['launches']} launches of one step"
            roof["algorithmic_bytes_per_launch"] = a["bytes"] / a["launches"]
        roof.update({"kernel": name, "share_of_step": round(a["ms"] / total_ms, 4), "launches": a["launches"],
                     "avg_launch_ms": a["ms"] / a["launches"], "algorithmic_bytes_per_step": a["bytes"],
                     "algorithmic_flops_per_step": a["flops"],
                     "roofline_ms_per_step": {"hbm": 1e3 * t_hbm, "tensor": 1e3 * t_tensor},
                     "peak_source": pk["source"] + " (MEASURED_PEAKS.json; sustained bf16 figure: kernel timed inside a long step)",
                     "timing": "CUDA events around each launch on the launching stream, separate instrumented step"})
    if world > 1:
        dist.barrier()
        if not embedded:
            dist.destroy_process_group()
    nparam, ws_total = tr.flat_p.numel(), tr.ws.bytes()
    del tr, net
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    cb = None
    if not args.no_cpu_baseline and not embedded:
        cb, _ = cpu_train_baseline(model, sample_shape)
    total_units = units * world * args.steps
    line = {
        "metric": METRIC[workload], "value": total_units / (ms_dev * 1e-3), "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16"}[precision], "data": "synthetic",
        "config": {"workload": workload, "model": model, "shape_per_gpu": list(shape), "global_batch": shape[0] * world,
                   "precision": precision,
                   "step": "forward + L1(clamp(out,0,1), clean) + hand-written backward + AdamW(lr 2e-4, wd 1e-2)"
                           + (" + NCCL all-reduce (sum, mean folded into AdamW) of the flat gradient buffer" if world > 1 else ""),
                   "task_id": "degradation id per sample, [B,1]",
                   "degradation": "one per sample of {Gaussian sigma~U(30,70)/255, non-iid per-band sigma {10,30,50,70}/255, random mask 0.7-0.9, band loss 0.1-0.3}",
                   "weights": "random-init (name-seeded synthetic), reference architecture, output conv x0.05",
                   "parallelism": f"dp{world} (batch-sharded, {nparam * 4 / 1e6:.1f} MB gradient all-reduce)" if world > 1 else "dp1",
                   "l2": "per-step working set (saved activations, GBs) exceeds the 126 MB L2; no explicit flush",
                   "cuda_graph": use_graph, "graph_ms": graph_ms, "losses_first_last": [losses[0], losses[-1]], "losses": [round(l, 5) for l in losses[:-1] + losses_e2e], "workspace_bytes": ws_total},
        "e2e": {"value": total_units / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": clean_h.numel() * 4 + h2d_extra[0], "d2h_bytes_per_step": 4,
                "note": "clean batch + per-sample degradation parameters travel; the degraded batch is synthesised on the device"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cb, "kernels": breakdown,
    }
    if embedded:
        return line
    print(json.dumps(line))
    return line


def inference_roofline(agg, passes: int, precision: str, workload: str):
    """dominant kernel family of an instrumented forward -> (roofline object, per-kernel breakdown)"""
    pk = peaks()
    total_ms = sum(a["ms"] for a in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    breakdown = {k: {"ms_per_step": round(a["ms"] / passes, 3), "share": round(a["ms"] / total_ms, 4),
                     "launches_per_step": a["launches"] // passes,
                     "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2) if a["ms"] else 0,
                     "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] else 0}
                 for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    # kernel FAMILIES: every tag is one __global__ template; the tcgen05 GEMM engine (all prologue/epilogue
    # variants of gemm_tc_kernel, incl. the implicit-GEMM convs) is reported as one kernel
    def family(tag):
        return "gemm_tc_kernel" if tag.startswith(("gemm_tc", "conv3x3_tc")) else tag
    fam = {}
    for k, a in agg.items():
        f = fam.setdefault(family(k), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for key in f:
            f[key] += a[key]
    name, a = max(fam.items(), key=lambda kv: kv[1]["ms"])
    sec = a["ms"] * 1e-3
    tensor_flops = a["flops"] * (3.0 if (name == "gemm_tc_kernel" and precision == "fp32") else 1.0)
    t_hbm = a["bytes"] / (pk["hbm_gbs"] * 1e9)
    t_tensor = tensor_flops / (pk["tf_sustained"] * 1e12)
    if t_tensor > t_hbm:
        ach = tensor_flops / sec / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sustained"], "traffic": None}
    else:
        ach = a["bytes"] / sec / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": None}
    tr_ = ncu_traffic(workload, name) if precision == "fp32" else None
    if tr_ is not None:
        roof["traffic"] = tr_["bytes_per_launch"]
        roof["traffic_source"] = f"{tr_['source']}: ncu dram read+write bytes / launch over the {tr_['launches']} launches of one step"
        roof["algorithmic_bytes_per_launch"] = a["bytes"] / a["launches"]
    roof.update({"kernel": name, "share_of_step": round(a["ms"] / total_ms, 4), "launches": a["launches"] // passes,
                 "avg_launch_ms": a["ms"] / a["launches"],
                 "algorithmic_bytes_per_step": a["bytes"] / passes, "algorithmic_flops_per_step": a["flops"] / passes,
                 "bf16_mma_flops_per_step": tensor_flops / passes,
                 "roofline_ms_per_step": {"hbm": 1e3 * t_hbm / passes, "tensor": 1e3 * t_tensor / passes},
                 "peak_source": pk["source"] + " (MEASURED_PEAKS.json; sustained bf16 figure: kernel timed inside a long step)",
                 "timing": "CUDA events around each launch on the launching stream, separate instrumented pass",
                 "note": "fp32 mode issues 3 bf16 MMAs per product (hi*hi+hi*lo+lo*hi); bytes = fp32 operands "
                         "read/written once (SURVEY 8d materialise-once model)"})
    return roof, breakdown


def run_sharded(args, embedded: bool = False):
    """`--workload cube512 --shard rows`: ONE 31x512x512 scene per step, its rows split over all ranks (BASELINE config 3 as
    specified; mp_hsir_b200/sharded.py) — strong scaling: value = 1 / (time per scene).  Collectives on the data path:
    neighbour halo send/recv before every block / conv, one all-reduce of the Gram statistics per block."""
    import torch.distributed as dist
    from mp_hsir_b200 import lib
    from mp_hsir_b200.parallel import max_over_ranks as _max
    from mp_hsir_b200.sharded import NcclComm, PeerComm, ShardedEngine, ThreadComm, _ThreadWorld, band_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1 and not embedded:
        dist.init_process_group("nccl", device_id=device)
    model, shape, unit, units, _, _ = WORKLOADS[args.workload]
    precision = args.precision
    cfg, net = build_net(model, device)
    net.set_precision(precision)
    if world == 1:
        comm = ThreadComm(_ThreadWorld(1), 0)
    elif args.comm == "peer":
        comm = PeerComm(device)          # fails loudly if the GPUs cannot map each other's memory; --comm nccl is the alternative
    else:
        comm = NcclComm()
    eng = ShardedEngine(net, comm)
    # graph replay needs collectives that are plain kernels on the capture stream (the peer-memory ones); capturing torch's
    # NCCL send/recv batches hung on this stack (NCCL 2.28.9, measured once, never again)
    if args.cuda_graph == "on" and isinstance(comm, NcclComm):
        raise SystemExit("--cuda-graph on with --shard rows needs --comm peer (NCCL calls are not captured)")
    eng.use_cuda_graph = args.cuda_graph != "off" and not isinstance(comm, NcclComm)   # auto = on: at 4-8 GPUs eager launches bound the band
    H = shape[2]
    x_full, clean, tid_host = make_input(shape, 0, args.workload)        # every rank derives the SAME scene from the seed
    r0, r1 = band_rows(H, rank, world)
    x_host = x_full[:, :, r0:r1].contiguous().pin_memory()
    x_dev = x_host.to(device)
    out_host = torch.empty_like(x_host).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            y = eng.forward_band(x_dev, tid_host, H)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        n0 = lib.LAUNCHES
        barrier()
        eng.use_cuda_graph, ug = False, eng.use_cuda_graph
        c0 = (comm.halo_exchanges, comm.all_reduces)
        eng.forward_band(x_dev, tid_host, H)       # one eager pass counts the collectives of a scene (a replay calls none from Python)
        comm_per_step = (comm.halo_exchanges - c0[0], comm.all_reduces - c0[1])
        eng.use_cuda_graph = ug
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = eng.forward_band(x_dev, tid_host, H)
        e1.record()
        barrier()
        launches = lib.LAUNCHES - n0
        ms_dev = _max(e0.elapsed_time(e1), device)
        # parity: assemble the scene on rank 0 and hold it against the unmodified reference's output
        if world > 1:
            parts = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
            dist.gather(y, parts, dst=0)
            y_full = torch.cat(parts, dim=2) if rank == 0 else None
        else:
            y_full = y
        parity = golden_check(args.workload, y_full, clean, precision) if rank == 0 else None
        # e2e: every rank's band comes from pinned host memory and goes back to it, inside the timed region
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for _ in range(args.steps):
            out_host.copy_(eng.forward_band(x_host.to(device, non_blocking=True), tid_host, H), non_blocking=True)
        f1.record()
        barrier()
        ms_e2e = _max(f0.elapsed_time(f1), device)
        clocks = sampler.stop() if sampler else None
        roof = breakdown = None
        if not args.no_roofline:
            # the instrumented pass contains collectives: every rank takes it, rank 0 with per-launch events
            if rank == 0:
                lib.PROFILER = lib.Profiler()
            eng.forward_band(x_dev, tid_host, H)
            if rank == 0:
                agg = lib.PROFILER.summary()
                lib.PROFILER = None
                roof, breakdown = inference_roofline(agg, 1, precision, args.workload)
                roof["traffic"] = None   # the committed ncu traffic capture is of the unsharded launch shapes
            barrier()
    ws_bytes, used_graph = eng.ws.bytes(), eng.use_cuda_graph
    if world > 1:
        dist.barrier()
        if hasattr(comm, "close"):
            comm.close()
        if not embedded:
            dist.destroy_process_group()
    del eng, net
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    cb = None
    if not args.no_cpu_baseline and not embedded:
        cb, _ = cpu_baseline(model, shape, float(units), unit, workload=args.workload)
    line = {
        "metric": METRIC[args.workload], "value": args.steps / (ms_dev * 1e-3), "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16"}[precision], "data": "synthetic",
        "config": {"workload": args.workload, "shard": "rows", "model": model, "scene": list(shape),
                   "rows_per_gpu": (r1 - r0), "halo_rows": 8, "task_id": tid_host.tolist(), "precision": precision,
                   "parallelism": f"ONE scene per step, rows split over {world} GPU(s): {comm_per_step[0]} neighbour halo exchanges "
                                  f"(send/recv, 8 rows each way) + {comm_per_step[1]} all-reduces of the Gram statistics "
                                  f"(<= 9.2 KB) per scene; exact (equals the single-GPU forward up to summation order)",
                   "collectives": comm.name,
                   "weights": "random-init (name-seeded synthetic), reference architecture",
                   "l2": "per-step working set exceeds the 126 MB L2 (GBs of activations); no explicit flush",
                   "cuda_graph": used_graph, "output_check": parity, "workspace_bytes_per_gpu": ws_bytes},
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": x_full.numel() * 4 + tid_host.numel() * 8, "d2h_bytes_per_step": x_full.numel() * 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cb, "kernels": breakdown,
    }
    if embedded:
        line.pop("kernels", None)
        line.pop("cpu_baseline", None)
        return line
    print(json.dumps(line))
    return line


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, shape, unit, units, sample_shape, frac = WORKLOADS[args.workload]
    if args.workload == "train64":
        cb, t = cpu_train_baseline(model, sample_shape, steps=args.steps, warmup=min(args.warmup, 1))
        steps_run = args.steps
    else:
        # the SAME config as our arm: the full workload shape, one whole batch per step (no extrapolation); the run is
        # bounded in time instead (the reference needs ~15-30 s per 512x512 cube on the host cores)
        cb, t = cpu_baseline(model, shape, float(units), unit, steps=args.steps, warmup=min(args.warmup, 1),
                             workload=args.workload, budget_s=args.reference_budget_s)
        sample_shape, frac, steps_run = shape, float(units), cb["steps_timed"]
    line = {"impl": "reference", "metric": METRIC[args.workload], "value": cb["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": steps_run, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t / frac * units,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "model": model, "shape_per_gpu": list(shape), "sample_shape": list(sample_shape),
                       "steps_requested": args.steps},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "cube512" and not args.no_train:
        tm, ts, tu, tn, tss, _ = WORKLOADS["train64"]
        tcb, tt = cpu_train_baseline(tm, tss, steps=1, warmup=0)
        line["train"] = {"impl": "reference", "metric": METRIC["train64"], "value": tcb["value"], "unit": tu,
                         "ms_per_step": 1e3 * tt / tss[0] * tn, "cpu_baseline": tcb}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cube512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="cube512 only: skip the embedded train64 measurement")
    ap.add_argument("--no-rows", action="store_true", help="N > 1, cube512: skip the embedded one-scene-over-all-GPUs measurement")
    ap.add_argument("--shard", default="cubes", choices=["cubes", "rows"],
                    help="cube512 with N GPUs: 'cubes' = one independent cube per GPU per step (weak scaling, default); "
                         "'rows' = ONE scene per step, its rows split over the GPUs with halo exchange + Gram all-reduce "
                         "(strong scaling, BASELINE config 3 as specified)")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="--shard rows: 'peer' = libmphsir one-kernel halo exchange / all-reduce over NVLink peer memory (CUDA IPC "
                         "windows), 'nccl' = torch.distributed send/recv + all_reduce")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch of the tcgen05 kernels (A/B switch; measured: no gain)")
    ap.add_argument("--reference-budget-s", type=float, default=120.0,
                    help="--impl reference: stop timing further steps once this many seconds of timed work are spent")
    ap.add_argument("--precision", default=None, choices=["fp32", "fp32_exact", "bf16"],
                    help="fp32 = tcgen05 with bf16 hi/lo split operands (meets the 1e-4 fp32 parity bound, default); "
                         "bf16 = bf16 operands (1e-2 bound); fp32_exact = FFMA")
    ap.add_argument("--cuda-graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the forward as one CUDA graph (auto: on for the launch-bound patch16 workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.precision is None:
        # inference is quoted at the fp32 parity bound; training follows the reference's mixed-precision trainer
        # (train.py:118 precision="16-mixed"; BASELINE config 4: bf16)
        args.precision = "bf16" if args.workload == "train64" else "fp32"
    if args.impl == "reference":
        return run_reference(args)
    if args.pdl:
        from mp_hsir_b200 import lib as _lib
        _lib.load().mphsir_debug_pdl(1)
    if args.workload == "train64":
        return run_train(args)
    if args.shard == "rows":
        if args.workload != "cube512" or args.precision == "fp32_exact":
            raise SystemExit("--shard rows applies to --workload cube512 with a tensor-core precision")
        return run_sharded(args)

    import torch.distributed as dist
    from mp_hsir_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    # BASELINE.json's metric has two halves: "HSI cubes/s (31x512x512 infer) & train patches/s".  The default run reports
    # the second half as a sub-object of the same line (same contract fields, its own roofline / e2e / cpu_baseline); it is
    # measured first, in a pristine allocator state, and everything it allocated is released before the inference part.
    train_line = None
    if args.workload == "cube512" and not args.no_train:
        train_line = run_train(args, embedded=True)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    model, shape, unit, units, sample_shape, frac = WORKLOADS[args.workload]
    cfg, net = build_net(model, device)
    net.set_precision(args.precision)
    precision = args.precision
    use_graph = args.cuda_graph == "on" or (args.cuda_graph == "auto" and args.workload == "patch16")
    net.use_cuda_graph = use_graph
    # small workloads fit the 126 MB L2: flush it between timed steps (per-step events); the 512^2 cube does not
    flush = args.workload != "cube512"
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if flush else None
    # independent cubes per rank (weak scaling): each rank gets its own seed
    x_host, clean_host, tid_host = make_input(shape, rank, args.workload)
    x_host, tid_host = x_host.pin_memory(), tid_host.pin_memory()
    x_dev = x_host.to(device)
    tid_dev = tid_host.to(device)
    out_host = torch.empty_like(x_host).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from mp_hsir_b200.parallel import max_over_ranks as _max

    def max_over_ranks(ms: float) -> float:
        return _max(ms, device)

    with torch.no_grad():
        for _ in range(args.warmup):
            y = net(x_dev, tid_dev)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        # ---- timed region: device-resident inputs ----------------------------------------------
        n0 = lib.LAUNCHES
        barrier()
        if flush:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in evs:
                flush_buf.fill_(1)
                a.record()
                y = net(x_dev, tid_dev)
                b.record()
            barrier()
            ms_local = sum(a.elapsed_time(b) for a, b in evs)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                y = net(x_dev, tid_dev)
            e1.record()
            barrier()
            ms_local = e0.elapsed_time(e1)
        launches = lib.LAUNCHES - n0
        ms_dev = max_over_ranks(ms_local)
        # a fast wrong answer is not a result: rank 0 ran the golden fixture's input, so the output of the timed loop must
        # match what the unmodified reference computed for it; the other ranks (other seeds) are checked for finiteness
        finite = bool(torch.isfinite(y).all())
        if not finite:
            raise SystemExit("bench: non-finite output")
        parity = golden_check(args.workload, y, clean_host, precision) if rank == 0 else None
        # ---- e2e: host buffers, H2D + D2H inside the timed region --------------------------------
        # the public host-to-host call: HostPipeline.restore_stream (pinned host cubes in, pinned host cubes out; the PCIe
        # copies of neighbouring cubes overlap the forward of the current one)
        from mp_hsir_b200.pipeline import HostPipeline
        pipe = HostPipeline(net, device)
        outs_host = [torch.empty_like(x_host).pin_memory() for _ in range(2)]
        pipe.restore_stream([(x_host, tid_host)] * 2, outs_host)
        barrier()
        e2e_check = float((outs_host[1].to(device) - y).abs().max())
        if e2e_check != 0.0:
            raise SystemExit(f"bench: pipelined host-to-host result differs from the device-resident one (max|d| = {e2e_check})")
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outs_all = [outs_host[i % 2] for i in range(args.steps)]
        f0.record()
        pipe.restore_stream([(x_host, tid_host)] * args.steps, outs_all)
        pipe.d2h.synchronize()
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        clocks = sampler.stop() if sampler else None

        # ---- roofline: instrumented pass (outside the timed region) -----------------------------
        roof, breakdown = None, None
        if rank == 0 and not args.no_roofline:
            lib.PROFILER = lib.Profiler()
            for _ in range(min(args.steps, 3)):
                net(x_dev, tid_dev)
            agg = lib.PROFILER.summary()
            lib.PROFILER = None
            roof, breakdown = inference_roofline(agg, min(args.steps, 3), precision, args.workload)

    ws_bytes = net.engine().ws.bytes()
    # N > 1: the same GPUs then restore ONE scene together (BASELINE config 3 as specified: rows split with halos, strong
    # scaling) — reported as the "scene_sharded" sub-object of the line, like "train"
    rows_line = None
    if world > 1 and args.workload == "cube512" and not args.no_rows and args.precision != "fp32_exact":
        del net
        torch.cuda.empty_cache()
        try:
            rows_line = run_sharded(args, embedded=True)
        except Exception as e:  # noqa: BLE001 - the weak-scaling line must not be lost to a transport problem
            rows_line = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cb = None
    if not args.no_cpu_baseline:
        cb, _ = cpu_baseline(model, shape, float(units), unit, workload=args.workload)   # ONE whole step of this workload
    total_units = units * world * args.steps
    line = {
        "metric": METRIC[args.workload], "value": total_units / (ms_dev * 1e-3), "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "fp32_exact": "f32", "bf16": "bf16"}[precision], "data": "synthetic",
        "config": {"workload": args.workload, "model": model, "shape_per_gpu": list(shape), "task_id": tid_host.tolist(), "precision": precision,
                   "precision_detail": {"fp32": "fp32 storage/accumulate; tensor-core products on bf16 hi+lo split operands (hi*hi+hi*lo+lo*hi), parity max|d|/max|ref| 2-3e-5 vs the fp32 reference (bound 1e-4)", "fp32_exact": "FFMA fp32", "bf16": "bf16 operands, fp32 accumulate/storage, parity 7e-3 (bound 1e-2)"}[precision],
                   "weights": "random-init (name-seeded synthetic), reference architecture",
                   "parallelism": f"independent cubes x{world} (no data-path collective)",
                   "l2": ("256 MB buffer written between timed steps (per-step CUDA events)" if flush else
                          "per-step working set (activations, GBs at 512x512) exceeds the 126 MB L2; no explicit flush"),
                   "cuda_graph": use_graph, "output_check": parity,
                   "workspace_bytes": ws_bytes},
        "e2e": {"value": total_units / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": x_host.numel() * 4 + tid_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cb, "kernels": breakdown,
    }
    if rows_line is not None:
        line["scene_sharded"] = rows_line
    if train_line is not None:
        train_line.pop("kernels", None)
        if not args.no_cpu_baseline:
            tm, _, _, _, tss, _ = WORKLOADS["train64"]
            train_line["cpu_baseline"], _ = cpu_train_baseline(tm, tss)
        line["train"] = train_line
    print(json.dumps(line))


if __name__ == "__main__":
    main()
