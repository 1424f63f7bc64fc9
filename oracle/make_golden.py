"""Generate tests/golden/* by running the UNMODIFIED reference (authoring container only).

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  Imports /root/reference/net/MP_HSIR.py through ``oracle.ref_import``
(timm/clip stubs), fills it with ``mp_hsir_b200.synth`` weights (seed 0), runs the cases
below under ``torch.no_grad()`` in eval mode on CPU/fp32 and stores inputs' *recipe* and the
outputs.  Inputs and weights are regenerated from their seeds by the tests, so only
outputs (and a few intermediates captured with forward hooks) are stored.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mp_hsir_b200.config import NetConfig  # noqa: E402
from mp_hsir_b200.synth import synthetic_input, synthetic_scene  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (model, input shape, task_id)
CASES = {
    "nat_b1_64": ("natural", (1, 31, 64, 64), [0]),                     # BASELINE config 1
    "nat_b4_64_mixed": ("natural", (4, 31, 64, 64), [0, 1, 2, 3]),      # TVSP cross-sample rows
    "nat_b2_64_task2d": ("natural", (2, 31, 64, 64), [[1], [4]]),       # [B,1] ids (train collate)
    "nat_b1_96x128": ("natural", (1, 31, 96, 128), [5]),                # non-construction, non-square
    "rs_b1_64": ("remote_sensing", (1, 100, 64, 64), [5]),              # wide spectral path
}

# BASELINE.json configs 2, 3 and 5 at their stated shapes.  The outputs are large, so a strided subsample
# (every `stride`-th pixel of every band) is stored together with per-band sums over ALL pixels (float64)
# and the reference's PSNR against the clean cube: name -> (model, input recipe, task_id, stride)
BIG_CASES = {
    "nat_b16_64": ("natural", ("rand", (16, 31, 64, 64)), [i % 6 for i in range(16)], 2),   # config 2
    "nat_cube512": ("natural", ("scene", 31, 512), [0], 4),                                  # config 3 (test.py:157-170)
    "rs_b1_256": ("remote_sensing", ("rand", (1, 100, 256, 256)), [3], 4),                   # config 5
}

TAP_CASE = ("natural", (1, 31, 32, 32), [2])
TAP_MODULES = {
    "encoder_level1.blocks.1.attn": "b1_attn",                          # Spatial_Attention, shifted
    "encoder_level1.blocks.1.local_spectral_attn": "b1_local",
    "encoder_level1.blocks.1.gobal_spectral_attn": "b1_global",
    "encoder_level1.blocks.1.mlp": "b1_mlp",
    "encoder_level1.blocks.1": "b1_out",
    "encoder_level1": "e1",
    "latent": "latent",
    "prompt1": "prompt1",
    "prompt2": "prompt2",
    "fusion1": "fusion1",
    "fusion2": "fusion2",
    "up2_1": "up2_1",
    "down1_2": "down1_2",
}


def cfg_of(model: str) -> NetConfig:
    return NetConfig.natural() if model == "natural" else NetConfig.remote_sensing()


def big_input(recipe):
    """(input, clean or None) of a BIG_CASES recipe — shared with tests/helpers.py through synth.py seeds."""
    if recipe[0] == "rand":
        return synthetic_input(tuple(recipe[1]), seed=0), None
    noisy, clean = synthetic_scene(recipe[1], recipe[2], seed=0)
    return noisy, clean


def psnr_per_band(y, clean):
    """utils/val_utils.py:49-69 of the reference: per-band PSNR, data_range 1, on clip(.,0,1), mean over bands."""
    y, c = y.clamp(0, 1).double(), clean.clamp(0, 1).double()
    mse = ((y - c) ** 2).mean(dim=(-1, -2))
    return float((10.0 * torch.log10(1.0 / mse)).mean())


def main_big():
    """python oracle/make_golden.py --big   (the 512x512 cube needs ~10 GB and about a minute on 8 cores)"""
    import time
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        meta = json.load(f)
    nets = {}
    for name, (model, recipe, tid, stride) in BIG_CASES.items():
        if model not in nets:
            nets[model] = ref_import.build_reference(cfg_of(model), seed=0)
        x, clean = big_input(recipe)
        t0 = time.time()
        with torch.no_grad():
            y = nets[model](x, torch.tensor(tid))
        dt = time.time() - t0
        sub = y[:, :, ::stride, ::stride].contiguous().numpy()
        band_sums = y.double().sum(dim=(-1, -2)).numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), sub=sub, band_sums=band_sums)
        meta[name] = {"model": model, "recipe": list(recipe), "shape": list(x.shape), "task_id": tid, "seed": 0,
                      "stride": stride, "out_absmax": float(y.abs().max()), "big": True}
        if clean is not None:
            meta[name]["psnr_ref_vs_clean"] = psnr_per_band(y, clean)
            meta[name]["psnr_in_vs_clean"] = psnr_per_band(x, clean)
        print(name, tuple(y.shape), float(y.abs().max()), f"{dt:.1f}s", meta[name].get("psnr_ref_vs_clean"), flush=True)
    with open(os.path.join(GOLDEN, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    nets = {}
    manifest = {}
    for model in ("natural", "remote_sensing"):
        net = ref_import.build_reference(cfg_of(model), seed=0)
        nets[model] = net
        params = {k for k, _ in net.named_parameters()}
        manifest[model] = [
            {"key": k, "shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", ""),
             "kind": "param" if k in params else "buffer"}
            for k, v in net.state_dict().items()
        ]
    with open(os.path.join(GOLDEN, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)

    meta = {}
    if os.path.exists(os.path.join(GOLDEN, "cases.json")):      # keep the --big entries (generated separately)
        with open(os.path.join(GOLDEN, "cases.json")) as f:
            meta = {k: v for k, v in json.load(f).items() if v.get("big")}
    for name, (model, shape, tid) in CASES.items():
        x = synthetic_input(shape, seed=0)
        with torch.no_grad():
            y = nets[model](x, torch.tensor(tid))
        np.savez(os.path.join(GOLDEN, name + ".npz"), out=y.numpy())
        meta[name] = {"model": model, "shape": list(shape), "task_id": tid, "seed": 0,
                      "out_absmax": float(y.abs().max())}
        print(name, tuple(y.shape), float(y.abs().max()))

    model, shape, tid = TAP_CASE
    net = nets[model]
    taps, hooks = {}, []
    mods = dict(net.named_modules())
    for mname, key in TAP_MODULES.items():
        def hook(_m, _i, o, key=key):
            taps[key] = o.detach().numpy().copy()
        hooks.append(mods[mname].register_forward_hook(hook))
    x = synthetic_input(shape, seed=0)
    with torch.no_grad():
        y = net(x, torch.tensor(tid))
    for h in hooks:
        h.remove()
    taps["out"] = y.numpy()
    np.savez(os.path.join(GOLDEN, "nat_b1_32_taps.npz"), **taps)
    meta["nat_b1_32_taps"] = {"model": model, "shape": list(shape), "task_id": tid, "seed": 0,
                              "taps": sorted(taps)}
    with open(os.path.join(GOLDEN, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", GOLDEN)


if __name__ == "__main__":
    main_big() if "--big" in sys.argv[1:] else main()
