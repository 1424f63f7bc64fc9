"""Generate tests/golden/nat_b2_32_grads.json: parameter gradients of the UNMODIFIED reference in train mode
(authoring container only; TEST INFRASTRUCTURE).

    python oracle/make_golden_grads.py

Case: natural model, synthetic weights (seed 0), x = synthetic_input((2,31,32,32)), task ids [[1],[4]] (the
[B,1] shape of the training collate), train mode with DropPath multipliers drawn once from a seeded generator
and injected into the timm stub in call order, objective  sum(out * R)  with R = seeded randn / numel (drives
every output element, unlike clamp+L1 which zeroes most of them on random weights).  Stored per parameter:
shape, L2 norm, and the dot product with a seeded random probe — enough to pin a full gradient tensor to
~1e-5 without shipping 58 MB.  ``tests/test_train_oracle.py`` checks the oracle's autograd against these; the
GPU tests compare the CUDA backward with the oracle element-wise.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mp_hsir_b200.config import NetConfig  # noqa: E402
from mp_hsir_b200.synth import synthetic_input, _gen  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHAPE = (2, 31, 32, 32)
TASK = [[1], [4]]


def keep_multipliers(cfg: NetConfig, B: int, seed: int = 7):
    """{(stage, block): [2,B] mask/keep_prob} for every block with a non-zero DropPath rate (net/MP_HSIR.py:780-805)."""
    g = _gen(seed, "droppath")
    keep = {}
    for st in cfg.stages():
        for i, rate in enumerate(st.dpr):
            if rate > 0.0:
                kp = 1.0 - rate
                keep[(st.name, i)] = (torch.rand(2, B, generator=g) < kp).float() / kp
    # make sure at least one sample is dropped somewhere so the path is exercised
    keep[("latent", 5)][0, 0] = 0.0
    keep[("refinement", 3)][1, 1] = 0.0
    return keep


def objective_weights(shape, seed: int = 11):
    g = _gen(seed, "objective")
    n = 1
    for s in shape:
        n *= s
    return torch.randn(shape, generator=g) / n


def probe(name: str, shape):
    return torch.randn(tuple(shape), generator=_gen(3, "probe/" + name))


FORWARD_ORDER = ["encoder_level1", "encoder_level2", "latent", "decoder_level2", "decoder_level1", "refinement"]


def main():
    cfg = NetConfig.natural()
    net = ref_import.build_reference(cfg, seed=0).train()
    x = synthetic_input(SHAPE, seed=0)
    keep = keep_multipliers(cfg, SHAPE[0])
    queue = []
    for name in FORWARD_ORDER:
        st = {s.name: s for s in cfg.stages()}[name]
        for i in range(st.depth):
            if (name, i) in keep:
                queue += [keep[(name, i)][0], keep[(name, i)][1]]
    ref_import.KEEP_QUEUE = queue
    out = net(x, torch.tensor(TASK))
    assert not queue, "DropPath call order mismatch"
    ref_import.KEEP_QUEUE = None
    # the training objective itself (train.py:58-61) on the same forward: value only (its gradient is mostly clamped away on
    # random weights, which is why the gradient fixture uses the linear objective below)
    clean = synthetic_input(SHAPE, seed=1)
    loss_l1 = float(torch.nn.L1Loss()(torch.clamp(out.detach(), 0, 1), clean))
    R = objective_weights(SHAPE)
    (out * R).sum().backward()
    rec = {}
    for n, p in net.named_parameters():
        if p.grad is None:
            rec[n] = None
            continue
        g = p.grad.double()
        rec[n] = {"shape": list(p.shape), "norm": float(g.norm()), "dot": float((g * probe(n, p.shape).double()).sum())}
    with open(os.path.join(GOLDEN, "nat_b2_32_grads.json"), "w") as f:
        json.dump({"shape": list(SHAPE), "task_id": TASK, "out_absmax": float(out.detach().abs().max()),
                   "out_sum": float(out.detach().double().sum()), "loss_l1_clamp": loss_l1, "grads": rec}, f, indent=0)
    none = [n for n, v in rec.items() if v is None]
    print("params", len(rec), "without grad", none)


if __name__ == "__main__":
    main()
