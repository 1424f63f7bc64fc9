"""CPU: pieces of bench.py that do not need a GPU — the synthetic training batch (BASELINE config 4) and the contract
keys of the reference arm's JSON line (run on a tiny step count)."""
import json
import os
import subprocess
import sys

import torch

from tests.conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_train_batch_recipes_are_seeded_and_mixed():
    a = bench.make_train_batch((16, 31, 64, 64), 3, 6)
    b = bench.make_train_batch((16, 31, 64, 64), 3, 6)
    assert all(torch.equal(x, y) for x, y in zip(a, b))              # same rank seed -> same batch
    noisy, clean, tid = a
    assert tid.shape == (16, 1) and tid.dtype == torch.int64 and set(tid.view(-1).tolist()) <= {0, 1, 2, 3}
    assert float(clean.min()) >= 0.0 and float(clean.max()) <= 1.0
    for i in range(16):
        k, d = int(tid[i, 0]), noisy[i] - clean[i]
        if k == 2:      # random mask: surviving pixels are untouched, 70-90 % are zeroed
            kept = noisy[i] != 0
            assert torch.equal(noisy[i][kept], clean[i][kept]) and 0.05 < float(kept.float().mean()) < 0.35
        elif k == 3:    # band loss: whole bands zeroed, the others untouched
            lost = (noisy[i].abs().sum(dim=(1, 2)) == 0)
            assert 1 <= int(lost.sum()) <= 10 and torch.equal(noisy[i][~lost], clean[i][~lost])
        else:           # additive noise with sigma in the reference's ranges
            assert 5.0 / 255 < float(d.std()) < 75.0 / 255
    c = bench.make_train_batch((16, 31, 64, 64), 4, 6)
    assert not torch.equal(c[0], noisy)                               # another rank trains on another shard


def test_workload_table_names_the_baseline_configs():
    assert bench.WORKLOADS["cube512"][1] == (1, 31, 512, 512) and bench.WORKLOADS["train64"][1] == (32, 31, 64, 64)
    assert bench.WORKLOADS["patch16"][1] == (16, 31, 64, 64) and bench.WORKLOADS["rs256"][1] == (1, 100, 256, 256)
    assert set(bench.METRIC) == set(bench.WORKLOADS)


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "patch16",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] in ("reference", "port") and j["vs_baseline"] is None
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    # same config as our arm: the full workload shape, no extrapolation from a smaller sample
    assert j["config"]["sample_shape"] == j["config"]["shape_per_gpu"] == [16, 31, 64, 64]


def test_golden_check_accepts_the_reference_and_rejects_a_perturbed_output():
    """bench.golden_check (the in-run output check) on the config-2 fixture: the oracle's output passes at the fp32
    bound, a 1e-3 perturbation fails."""
    import pytest
    from mp_hsir_b200.config import NetConfig
    from mp_hsir_b200.synth import synthetic_clip_prompt
    from oracle import mp_hsir_oracle as O
    from tests.helpers import synthetic_state_dict
    x, clean, tid = bench.make_input((16, 31, 64, 64), 0, "patch16")
    assert tid.tolist() == [i % 6 for i in range(16)] and clean is None
    cfg = NetConfig.natural()
    with torch.no_grad():
        y = O.forward(synthetic_state_dict("natural"), cfg, x, tid, synthetic_clip_prompt(cfg.task_classes))
    res = bench.golden_check("patch16", y, None, "fp32")
    assert res["max_abs_err_over_max_abs_ref"] < 2e-5
    with pytest.raises(SystemExit):
        bench.golden_check("patch16", y + 1e-3 * y.abs().max(), None, "fp32")
