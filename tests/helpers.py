"""Shared builders for the parity tests (weights/inputs are regenerated from seeds)."""
import functools

import torch

from mp_hsir_b200.config import NetConfig
from mp_hsir_b200.synth import synth_tensor, synthetic_clip_prompt, synthetic_input
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
