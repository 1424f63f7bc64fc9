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
from mp_hsir_b200.synth import synthetic_input  # noqa: E402
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
    main()
