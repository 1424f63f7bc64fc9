"""Shared builders for the parity tests (weights/inputs are regenerated from seeds)."""
import functools

import torch

from mp_hsir_b200.config import NetConfig
from mp_hsir_b200.synth import synth_tensor, synthetic_clip_prompt, synthetic_input, synthetic_scene
from tests.conftest import GOLDEN  # noqa: F401

import json
import os


def cfg_of(model: str) -> NetConfig:
    return NetConfig.natural() if model == "natural" else NetConfig.remote_sensing()


@functools.lru_cache(maxsize=2)
def synthetic_state_dict(model: str, seed: int = 0):
    """state_dict the reference would hold after ``fill_state_dict_`` — rebuilt from the manifest,
    no reference import needed (runs on the GPU box)."""
    with open(os.path.join(GOLDEN, "state_dict_manifest.json")) as f:
        manifest = json.load(f)[model]
    sd = {}
    for e in manifest:
        if e["kind"] == "param":
            sd[e["key"]] = synth_tensor(e["key"], e["shape"], seed)
    return sd


def case_inputs(meta):
    x = synthetic_input(tuple(meta["shape"]), seed=meta["seed"])
    tid = torch.tensor(meta["task_id"])
    return x, tid


def clip_for(cfg: NetConfig):
    return synthetic_clip_prompt(cfg.task_classes)


def big_case_inputs(meta):
    """(input, clean-or-None, task ids) of a BASELINE-shape case (oracle/make_golden.py BIG_CASES)."""
    recipe = meta["recipe"]
    tid = torch.tensor(meta["task_id"])
    if recipe[0] == "rand":
        return synthetic_input(tuple(recipe[1]), seed=meta["seed"]), None, tid
    noisy, clean = synthetic_scene(recipe[1], recipe[2], seed=meta["seed"])
    return noisy, clean, tid


def psnr_per_band(y, clean) -> float:
    """utils/val_utils.py:49-69 of the reference: per-band PSNR (data_range 1) on clip(.,0,1), mean over bands and batch."""
    y, c = y.clamp(0, 1).double(), clean.clamp(0, 1).double()
    mse = ((y - c) ** 2).mean(dim=(-1, -2))
    return float((10.0 * torch.log10(1.0 / mse)).mean())


def big_case_errors(y: torch.Tensor, g: dict, meta: dict):
    """parity of a full-size output against the stored strided subsample + all-pixel band sums of the reference:
    (max|d|/max|ref| on the subsample, max band-mean error / max|ref|)."""
    s = meta["stride"]
    sub = y[:, :, ::s, ::s].double().cpu()
    e_sub = float((sub - g["sub"].double()).abs().max() / meta["out_absmax"])
    hw = y.shape[-1] * y.shape[-2]
    e_mean = float((y.double().sum(dim=(-1, -2)).cpu() - g["band_sums"].double()).abs().max() / hw / meta["out_absmax"])
    return e_sub, e_mean
