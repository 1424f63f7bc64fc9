"""Shared pieces of the training-path tests: the gradient case of tests/golden/nat_b2_32_grads.json rebuilt from
seeds (no reference import), and the oracle's autograd gradients for it."""
import functools
import json
import os

import torch

from mp_hsir_b200.config import NetConfig
from mp_hsir_b200.synth import _gen, synthetic_clip_prompt, synthetic_input
from oracle import mp_hsir_oracle as O
from tests.conftest import GOLDEN
from tests.helpers import synthetic_state_dict


def keep_multipliers(cfg: NetConfig, B: int, seed: int = 7):
    """Must stay identical to oracle/make_golden_grads.py:keep_multipliers."""
    g = _gen(seed, "droppath")
    keep = {}
    for st in cfg.stages():
        for i, rate in enumerate(st.dpr):
            if rate > 0.0:
                kp = 1.0 - rate
                keep[(st.name, i)] = (torch.rand(2, B, generator=g) < kp).float() / kp
    keep[("latent", 5)][0, 0] = 0.0
    keep[("refinement", 3)][1, 1] = 0.0
    return keep


def objective_weights(shape, seed: int = 11):
    g = _gen(seed, "objective")
    n = 1
    for s in shape:
        n *= s
    return torch.randn(tuple(shape), generator=g) / n


def probe(name: str, shape):
    return torch.randn(tuple(shape), generator=_gen(3, "probe/" + name))


def golden_grads():
    with open(os.path.join(GOLDEN, "nat_b2_32_grads.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=1)
def oracle_grads():
    """(out, {param name: grad}) of the oracle on the golden gradient case (CPU autograd, fp32)."""
    gold = golden_grads()
    cfg = NetConfig.natural()
    sd = {k: v.clone().requires_grad_(True) for k, v in synthetic_state_dict("natural").items()}
    x = synthetic_input(tuple(gold["shape"]), seed=0)
    tid = torch.tensor(gold["task_id"])
    keep = keep_multipliers(cfg, x.shape[0])
    out = O.forward(sd, cfg, x, tid, synthetic_clip_prompt(cfg.task_classes), keep=keep)
    (out * objective_weights(x.shape)).sum().backward()
    return out.detach(), {k: v.grad for k, v in sd.items()}
