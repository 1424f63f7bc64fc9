"""tests/golden/nat_b2_32_grads_bf16_yardstick.json: how far the UNMODIFIED reference's own bf16 mixed-precision gradients
are from its fp32 gradients, per parameter tensor (authoring container only; TEST INFRASTRUCTURE).

    python oracle/make_golden_bf16_yardstick.py

Same case as oracle/make_golden_grads.py (natural model, synthetic weights, B=2 32x32, [B,1] task ids, injected DropPath
multipliers, objective sum(out*R)).  The reference trains under Lightning ``precision="16-mixed"`` (train.py:118), i.e.
``torch.autocast``; here it runs once in fp32 and once under ``torch.autocast("cpu", dtype=torch.bfloat16)`` and the
relative L2 distance of every parameter gradient between the two runs is stored.  The bf16 mode of the CUDA trainer is
then held to "no worse than the reference's own autocast" per tensor (tests/test_train_model_gpu.py) instead of an
arbitrary tolerance.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mp_hsir_b200.config import NetConfig  # noqa: E402
from mp_hsir_b200.synth import synthetic_input  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.make_golden_grads import FORWARD_ORDER, SHAPE, TASK, keep_multipliers, objective_weights  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(autocast: bool):
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
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = net(x, torch.tensor(TASK))
    ref_import.KEEP_QUEUE = None
    (out.float() * objective_weights(SHAPE)).sum().backward()
    return out.detach().float(), {n: p.grad.detach().double() for n, p in net.named_parameters() if p.grad is not None}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out32, g32 = run(False)
    out16, g16 = run(True)
    rec = {n: float((g16[n] - g32[n]).norm() / g32[n].norm().clamp_min(1e-30)) for n in g32}
    vals = sorted(rec.values())
    summary = {"median": vals[len(vals) // 2], "p90": vals[int(0.9 * len(vals))], "max": vals[-1],
               "out_rel_err": float((out16 - out32).abs().max() / out32.abs().max())}
    with open(os.path.join(GOLDEN, "nat_b2_32_grads_bf16_yardstick.json"), "w") as f:
        json.dump({"what": "relative L2 distance of the unmodified reference's parameter gradients under "
                           "torch.autocast(cpu, bfloat16) from its fp32 gradients, same case as nat_b2_32_grads.json",
                   "summary": summary, "rel_l2": rec}, f, indent=0)
    print(len(rec), summary)


if __name__ == "__main__":
    main()
