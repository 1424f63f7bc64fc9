"""CPU: host-side logic of the drop-in module (no GPU, no compute through the library)."""
import json
import os

import pytest
import torch

from mp_hsir_b200 import MP_HSIR_Net, engine as E
from mp_hsir_b200.config import NetConfig
from mp_hsir_b200.synth import synth_tensor
from oracle import mp_hsir_oracle as O
from tests.conftest import GOLDEN


@pytest.mark.parametrize("model,args", [("natural", (31, 31, 64, 6)), ("remote_sensing", (100, 100, 96, 7))])
def test_state_dict_abi_matches_reference_manifest(model, args):
    """658 keys, same order / shapes / dtypes as the unmodified reference."""
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))[model]
    net = MP_HSIR_Net(args[0], args[1], args[2], task_classes=args[3])
    sd = net.state_dict()
    assert len(sd) == 658
    assert [e["key"] for e in man] == list(sd)
    for e in man:
        t = sd[e["key"]]
        assert list(t.shape) == e["shape"] and str(t.dtype) == "torch." + e["dtype"], e["key"]
    params = {k for k, _ in net.named_parameters()}
    assert {e["key"] for e in man if e["kind"] == "param"} == params
    # buffers hold the same values the reference computes at construction
    assert torch.equal(sd["encoder_level1.blocks.1.attn_mask"], O.shift_mask(64, 64))
    assert torch.equal(sd["latent.blocks.3.attn_mask"], O.shift_mask(16, 16))
    idx = sd["latent.blocks.0.attn.relative_position_index"]
    assert idx.dtype == torch.int64 and int(idx[0, 63]) == 0 and int(idx[63, 0]) == 224


def test_strict_load_of_a_lightning_style_checkpoint():
    net = MP_HSIR_Net()
    ckpt = {"state_dict": {"net." + k: v.clone() for k, v in net.state_dict().items()}}
    stripped = {k[4:]: v for k, v in ckpt["state_dict"].items()}
    missing, unexpected = MP_HSIR_Net().load_state_dict(stripped, strict=True)
    assert not missing and not unexpected


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="task_classes must be 6 or 7"):
        MP_HSIR_Net(task_classes=3)
    with pytest.raises(ValueError):
        MP_HSIR_Net(clip_prompt=torch.zeros(5, 512))
    MP_HSIR_Net(task_classes=1)


def test_no_cpu_fallback():
    net = MP_HSIR_Net().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 31, 32, 32), torch.tensor([0]))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 31, 32, 32))


def test_stage_table_matches_reference_hyperparameters():
    st = NetConfig.natural().stages()
    assert [(s.name, s.depth, s.dim, s.heads, s.rank) for s in st] == [
        ("encoder_level1", 2, 64, 2, 8), ("encoder_level2", 4, 128, 4, 8), ("latent", 6, 256, 8, 8),
        ("decoder_level2", 4, 128, 4, 8), ("decoder_level1", 2, 128, 2, 16), ("refinement", 4, 128, 2, 16)]
    rs = NetConfig.remote_sensing()
    assert [rs.hidden(d) for d in (96, 192, 384)] == [255, 510, 1021]
    dpr = NetConfig.natural().drop_path_rates()
    assert dpr[0] == 0.0 and abs(dpr[-1] - 0.1) < 1e-9 and len(dpr) == 12
    assert st[5].dpr == st[1].dpr  # refinement reuses dpr[2:6] (net/MP_HSIR.py:805)


def test_weight_packing_is_pure_relayout():
    w, b = torch.randn(340, 64), torch.randn(340)
    wt, bias = E.pack_glu_fc1(w, b, 170, 176)
    assert wt.shape == (64, 384) and bias.shape == (384,)
    assert torch.equal(wt[:, 0:340:2], w[:170].t()) and torch.equal(wt[:, 1:340:2], w[170:].t())
    assert wt[:, 340:].abs().sum() == 0 and torch.equal(bias[1:340:2], b[170:])
    c = torch.randn(32, 31, 3, 3)
    p = E.pack_conv3x3(c, cin_pad=32)
    assert p.shape == (9 * 32, 64) and torch.equal(p.view(9, 32, 64)[4, :31, :32], c[:, :, 1, 1].t())
    assert p.view(9, 32, 64)[:, 31].abs().sum() == 0
    u = torch.randn(64, 16, 3, 3)
    ps = E.pack_conv3x3(u, shuffle=True).view(9, 16, 64)
    # packed column q*Cn+cn holds reference channel cn*4+q
    assert torch.equal(ps[0, :, 1 * 16 + 5], u[5 * 4 + 1, :, 0, 0])
    pin, w9, pout = E.pack_gdfn(torch.randn(340, 64, 1, 1), torch.randn(340, 1, 3, 3), torch.randn(64, 170, 1, 1), 170, 176)
    assert pin.shape == (64, 384) and w9.shape == (9, 352) and pout.shape == (176, 64)
    assert w9[:, 170:176].abs().sum() == 0 and pout[170:].abs().sum() == 0


def test_synthetic_weights_are_name_seeded():
    a = synth_tensor("latent.blocks.0.attn.qkv.weight", (768, 256), 0)
    b = synth_tensor("latent.blocks.0.attn.qkv.weight", (768, 256), 0)
    c = synth_tensor("latent.blocks.1.attn.qkv.weight", (768, 256), 0)
    assert torch.equal(a, b) and not torch.equal(a, c)


def _gemm_plan(M, N, K, rows_per_batch=0, per_sample=0, sms=148):
    import ctypes
    from mp_hsir_b200 import lib
    out = (ctypes.c_int * 8)()
    assert lib.load().mphsir_gemm_plan(M, N, K, rows_per_batch, per_sample, sms, out) == 0
    return dict(zip(("cluster", "psplit", "ppg", "grid", "iters", "rev", "n_full", "pass_cols"), out))


@pytest.mark.parametrize("M,N,K", [(262144, 384, 128), (65536, 512, 128), (65536, 384, 128), (16384, 768, 256), (4096, 1376, 256),
                                   (4096, 256, 688), (8192, 1024, 256), (128, 64, 64), (300, 288, 96), (262144, 256, 1152),
                                   (4096, 512, 2304), (65536, 128, 128), (19000, 512, 64)])
def test_gemm_work_plan_covers_every_pass_exactly_once(M, N, K):
    """mphsir_gemm_plan (the host arithmetic of the tensor-core GEMM launcher, no device needed): every (row tile, 256-column
    pass) belongs to exactly one work item, the grid fits the machine, pairs are even, few-tile launches and half-empty last
    rounds are split."""
    pl = _gemm_plan(M, N, K)
    assert pl["pass_cols"] in (128, 256)
    tiles, npass = (M + 127) // 128, ((N + 15) // 16 * 16 + pl["pass_cols"] - 1) // pl["pass_cols"]
    assert pl["cluster"] in (1, 2) and pl["psplit"] >= 1 and pl["ppg"] >= 1 and 0 <= pl["n_full"] <= tiles
    # decode every work item exactly as the kernel does (gemm_tc.cu tc_decode) and count the (tile, pass) pairs
    seen = {}
    items = pl["n_full"] + (tiles - pl["n_full"]) * pl["psplit"]
    for w in range(items):
        if w < pl["n_full"]:
            t, p0, p1 = w, 0, npass
        else:
            idx = w - pl["n_full"]
            t = pl["n_full"] + idx // pl["psplit"]
            p0 = (idx % pl["psplit"]) * pl["ppg"]
            p1 = min(npass, p0 + pl["ppg"])
        assert p1 > p0
        for ps in range(p0, p1):
            seen[(t, ps)] = seen.get((t, ps), 0) + 1
    assert len(seen) == tiles * npass and set(seen.values()) == {1}
    assert 1 <= pl["grid"] <= 148 and pl["grid"] * pl["iters"] >= items > pl["grid"] * (pl["iters"] - 1) - (pl["cluster"] - 1)
    if pl["cluster"] == 2:
        assert pl["grid"] % 2 == 0 and pl["psplit"] == 1 and tiles >= 2
    if tiles * 2 <= 148 and N > 128:
        assert pl["psplit"] >= 2 and pl["n_full"] == 0, "few-tile launches hand the passes of a row tile to several CTAs"
        assert pl["pass_cols"] == (128 if N <= 256 else 256)
    if tiles > 148 and 0 < tiles % 148 <= 74 and npass >= 2:
        assert pl["psplit"] >= 2 and pl["n_full"] == tiles - tiles % 148, "a half-empty last round is split"
        assert pl["iters"] == tiles // 148 + 1
    assert pl["rev"] == (1 if M >= 131072 else 0)


def test_gemm_work_plan_pairs_need_shared_weights_within_a_pair():
    # per-sample weights: a CTA pair runs tiles (2q, 2q+1) on one instruction stream, so both must lie in the same sample
    assert _gemm_plan(3 * 384, 256, 1152, rows_per_batch=384, per_sample=1)["cluster"] == 1      # 3 tiles per sample: odd
    assert _gemm_plan(64 * 512, 256, 1152, rows_per_batch=512, per_sample=1)["cluster"] == 2     # 4 tiles per sample
