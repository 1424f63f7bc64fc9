"""CPU: host-side pieces of the training path — the per-epoch LR schedule against values produced by the unmodified
reference scheduler (tests/golden/lr_schedule.json, oracle/make_golden_lr.py) and the Lightning-format checkpoint
interchange (train.py:107-116, test.py:575)."""
import json
import os

import torch

from mp_hsir_b200 import MP_HSIR_Net
from mp_hsir_b200.checkpoint import load_reference_checkpoint, reference_state_dict, save_reference_checkpoint
from mp_hsir_b200.schedule import EpochSchedule, warmup_cosine_lr
from mp_hsir_b200.synth import fill_state_dict_
from tests.conftest import GOLDEN


def test_lr_schedule_matches_reference_scheduler():
    gold = json.load(open(os.path.join(GOLDEN, "lr_schedule.json")))
    for name, g in gold.items():
        for epoch, ref in enumerate(g["lrs"]):
            got = warmup_cosine_lr(epoch, g["base_lr"], g["warmup_epochs"], g["max_epochs"], 0.0, g["eta_min"])
            assert abs(got - ref) <= 1e-13 + 1e-6 * abs(ref), (name, epoch, got, ref)
    # first epoch trains at lr = 0 when stepped per epoch (the reference's documented caveat, utils/schedulers.py:242-245)
    assert warmup_cosine_lr(0, 2e-4, 50, 500) == 0.0


def test_epoch_schedule_drives_trainer_lr():
    class T:
        lr = -1.0
    t = T()
    s = EpochSchedule(t, base_lr=2e-4, max_epochs=100)
    assert s.warmup_epochs == 10
    assert s.begin_epoch(0) == 0.0 and t.lr == 0.0
    assert abs(s.begin_epoch(9) - 2e-4) < 1e-12
    assert abs(s.begin_epoch(10) - 2e-4) < 1e-12
    assert s.begin_epoch(99) < 2e-6


def test_reference_checkpoint_round_trip(tmp_path):
    net = MP_HSIR_Net(31, 31, 64, task_classes=6)
    fill_state_dict_(net, seed=3)
    path = str(tmp_path / "epoch=49.ckpt")
    save_reference_checkpoint(path, net, epoch=49, global_step=1234)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) >= {"state_dict", "epoch", "global_step"}
    assert all(k.startswith("net.") for k in ckpt["state_dict"]) and len(ckpt["state_dict"]) == 658
    other = MP_HSIR_Net(31, 31, 64, task_classes=6)
    rep = load_reference_checkpoint(path, other)
    assert len(rep["loaded"]) == 658 and not rep["skipped"]
    a, b = reference_state_dict(net), reference_state_dict(other)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_checkpoint_filter_skips_mismatched_shapes():
    """train.py:111-116: keys that are missing or whose shape differs are dropped, the rest load (strict=False)."""
    src = MP_HSIR_Net(31, 31, 64, task_classes=6)
    fill_state_dict_(src, seed=4)
    sd = reference_state_dict(src)
    sd["net.patch_embed.proj.weight"] = torch.zeros(64, 100, 3, 3)       # a 100-band checkpoint's stem
    sd["net.not_a_key"] = torch.zeros(3)
    dst = MP_HSIR_Net(31, 31, 64, task_classes=6)
    before = dst.patch_embed.proj.weight.detach().clone()
    rep = load_reference_checkpoint({"state_dict": sd}, dst)
    assert set(rep["skipped"]) == {"patch_embed.proj.weight", "not_a_key"} and len(rep["loaded"]) == 657
    assert torch.equal(dst.patch_embed.proj.weight, before)
    assert torch.equal(dst.output.weight, src.output.weight)
