"""GPU: device PSNR/SSIM (utils/val_utils.py:49-105) and device degradations against the CPU restatements."""
import numpy as np
import pytest
import torch

from mp_hsir_b200 import lib
from mp_hsir_b200 import metrics as GM
from mp_hsir_b200.degrade import degrade, degrade_batch, draw_parameters
from mp_hsir_b200.synth import synthetic_input, synthetic_scene
from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 5, 40, 50), (1, 31, 64, 64), (1, 3, 7, 7), (1, 2, 135, 263)])
def test_psnr_ssim_matches_oracle(shape):
    clean = synthetic_input(shape, seed=3)
    rec = clean + 0.08 * torch.randn(shape, generator=torch.Generator().manual_seed(1)) - 0.01   # leaves [0,1]: clip is live
    p, s, n = GM.compute_psnr_ssim(rec.cuda(), clean.cuda())
    pr, sr, nr = M.compute_psnr_ssim(rec.numpy(), clean.numpy())
    assert n == nr == shape[0]
    assert abs(p - pr) < 1e-9 * max(1.0, abs(pr)) and abs(s - sr) < 1e-9, (p, pr, s, sr)


def test_psnr_ssim_on_the_scene_shape_and_band_miss_variant():
    noisy, clean = synthetic_scene(31, 512, seed=0)
    p, s, _ = GM.compute_psnr_ssim(noisy.cuda(), clean.cuda())
    pr, sr, _ = M.compute_psnr_ssim(noisy.numpy(), clean.numpy())
    assert abs(p - pr) < 1e-8 and abs(s - sr) < 1e-9
    deg = clean.clone()
    deg[0, [2, 17, 30]] = 0
    p2, s2, n2 = GM.compute_psnr_ssim2(noisy.cuda(), clean.cuda(), deg.cuda())
    pr2, sr2, nr2 = M.compute_psnr_ssim2(noisy.numpy(), clean.numpy(), deg.numpy())
    assert n2 == nr2 == 1 and abs(p2 - pr2) < 1e-8 and abs(s2 - sr2) < 1e-9
    assert GM.compute_psnr_ssim2(noisy.cuda(), clean.cuda(), clean.cuda() + 1.0) == (0, 0, 0)   # nothing lost: nothing counted


def test_degrade_matches_oracle_stream():
    B, C, H, W = 3, 7, 24, 40
    clean = synthetic_input((B, C, H, W), seed=9)
    g = torch.Generator().manual_seed(5)
    sigma = 0.3 * torch.rand(B, C, generator=g)
    keep = (torch.rand(B, C, generator=g) > 0.3).float()
    ratio = torch.tensor([-1.0, 0.7, 0.9])
    out = degrade(clean.cuda(), sigma, keep, ratio, seed=0x1234567890ABCDEF).cpu().numpy()
    ref, um, n = M.degrade(clean.numpy(), sigma.numpy(), keep.numpy(), ratio.numpy(), seed=0x1234567890ABCDEF)
    # the mask decisions are integer arithmetic (bit exact); the normals go through fp32 log / sincospi
    assert np.array_equal(out == 0, ref == 0) or np.mean((out == 0) != (ref == 0)) < 1e-6
    assert np.max(np.abs(out - ref)) < 2e-6
    # another seed, another stream
    out2 = degrade(clean.cuda(), sigma, keep, ratio, seed=1).cpu().numpy()
    assert np.mean(out2 != out) > 0.5


def test_degrade_batch_follows_the_recipes():
    clean = synthetic_input((32, 31, 64, 64), seed=2).cuda()
    before = lib.LAUNCHES
    noisy, tid = degrade_batch(clean, seed=77, generator=torch.Generator().manual_seed(3))
    assert lib.LAUNCHES - before == 1 and tid.shape == (32, 1)
    d = (noisy - clean).cpu()
    c = clean.cpu()
    for b in range(32):
        k = int(tid[b, 0])
        if k == 0:
            assert 28 / 255 < float(d[b].std()) < 72 / 255
        elif k == 1:
            sd = d[b].std(dim=(1, 2)) * 255
            assert all(min(abs(float(v) - t) for t in (10, 30, 50, 70)) < 3 for v in sd)
        elif k == 2:
            kept = noisy[b].cpu() != 0
            assert 0.05 < float(kept.float().mean()) < 0.35 and torch.equal(noisy[b].cpu()[kept], c[b][kept])
        else:
            lost = noisy[b].abs().sum(dim=(1, 2)).cpu() == 0
            assert int(lost.sum()) in (3, 6, 9) and torch.equal(noisy[b].cpu()[~lost], c[b][~lost])


@pytest.mark.gpu
def test_gaussian_blur_matches_oracle():
    """mphsir_gaussian_blur vs the CPU restatement of utils/degradation_utils.py:91-108: every kernel size of the reference's
    de_dict, a non-square plane that is not a multiple of the 32 x 32 tile, samples with ksize 0 untouched"""
    from mp_hsir_b200.degrade import gaussian_blur
    clean = synthetic_input((6, 5, 72, 50), seed=4)
    ksize = torch.tensor([9, 0, 15, 21, 7, 11], dtype=torch.int32)
    out = torch.full_like(clean, 7.0).cuda()
    gaussian_blur(clean.cuda(), ksize, out=out)
    out = out.cpu().numpy()
    for b in range(6):
        k = int(ksize[b])
        if k == 0:
            assert (out[b] == 7.0).all()
        else:
            ref = M.gaussian_blur(clean[b].numpy(), k)
            assert abs(out[b] - ref).max() < 2e-6


@pytest.mark.gpu
def test_degrade_batch_with_blur_recipe():
    from mp_hsir_b200.degrade import ALL_RECIPES
    clean = synthetic_input((16, 31, 64, 64), seed=5).cuda()
    noisy, tid = degrade_batch(clean, seed=9, de_types=ALL_RECIPES[:5], generator=torch.Generator().manual_seed(6))
    assert 4 in tid.view(-1).tolist()
    for b in range(16):
        if int(tid[b, 0]) == 4:   # blurred: smoother than the clean patch, same mean to a few percent
            assert float(noisy[b].var()) < float(clean[b].var()) and abs(float(noisy[b].mean() - clean[b].mean())) < 0.05


@pytest.mark.gpu
def test_sr_degrade_matches_oracle():
    """mphsir_sr_degrade vs the CPU restatement of utils/degradation_utils.py:165-176 + :189-200: every factor of the reference's
    de_dict, a non-square plane, samples with factor 0 untouched; fp32 sums of 16 products vs float64: 2e-6"""
    from mp_hsir_b200.degrade import sr_degrade
    clean = synthetic_input((5, 4, 64, 96), seed=4)
    factor = torch.tensor([2, 0, 4, 8, 2], dtype=torch.int32)
    out = torch.full_like(clean, 7.0).cuda()
    sr_degrade(clean.cuda(), factor, out=out)
    out = out.cpu().numpy()
    for b in range(5):
        f = int(factor[b])
        if f == 0:
            assert (out[b] == 7.0).all()
        else:
            ref = M.sr_degrade(clean[b].numpy(), f)
            assert abs(out[b] - ref).max() < 2e-6
    with pytest.raises(ValueError):
        sr_degrade(clean.cuda(), torch.tensor([128, 0, 0, 0, 0], dtype=torch.int32))


@pytest.mark.gpu
def test_degrade_batch_reference_default_list():
    """the reference's default natural-scene recipe list (options.py:15), all six synthesised on the device; task id = position"""
    from mp_hsir_b200.degrade import REFERENCE_DEFAULT
    clean = synthetic_input((24, 31, 64, 64), seed=7).cuda()
    noisy, tid = degrade_batch(clean, seed=11, de_types=REFERENCE_DEFAULT, generator=torch.Generator().manual_seed(4))
    kinds = [REFERENCE_DEFAULT[int(t)] for t in tid.view(-1)]
    assert {"sr", "blur"} <= set(kinds)
    for b, kind in enumerate(kinds):
        assert not torch.equal(noisy[b], clean[b])
        if kind == "sr":   # piecewise constant on f x f blocks with f >= 2
            assert torch.equal(noisy[b][:, ::2, ::2], noisy[b][:, 1::2, 1::2])
        elif kind == "blur":
            assert float(noisy[b].var()) < float(clean[b].var())


@pytest.mark.gpu
def test_blur2d_matches_oracle():
    """mphsir_blur2d vs explicit shifted sums: the reference's circle kernel (utils/degradation_utils.py:110-128), a square kernel
    (:150-163) and an asymmetric random kernel (orientation: F.conv2d is a cross-correlation); inactive samples untouched"""
    from mp_hsir_b200.degrade import blur2d, circle_kernel
    clean = synthetic_input((3, 4, 72, 50), seed=9)
    g = torch.Generator().manual_seed(1)
    for kernel in (circle_kernel(9), torch.full((5, 5), 1.0 / 25.0), torch.rand(7, 7, generator=g), circle_kernel(21)):
        out = torch.full_like(clean, 7.0).cuda()
        blur2d(clean.cuda(), kernel, torch.tensor([1, 0, 1]), out=out)
        out = out.cpu().numpy()
        assert (out[1] == 7.0).all()
        for b in (0, 2):
            ref = M.blur2d(clean[b].numpy(), kernel.numpy())
            assert abs(out[b] - ref).max() < 1e-5 * max(1.0, float(kernel.sum()))
    with pytest.raises(ValueError):
        blur2d(clean.cuda(), torch.ones(4, 4))


@pytest.mark.gpu
def test_poisson_matches_oracle_stream():
    """mphsir_poisson vs the oracle's restatement of the same Philox stream and CDF inversion: the counts are integers decided
    by comparisons in float64, so they agree exactly (a last-bit difference of exp() could move at most an isolated element)"""
    from mp_hsir_b200.degrade import poisson_noise
    x = synthetic_input((4, 5, 24, 40), seed=2)
    x[3] = x[3] - 0.5                                           # negative inputs are clipped to lambda = 0
    scale = torch.tensor([10.0, 0.0, 4.0, 10.0])
    got = poisson_noise(x.cuda(), scale, seed=0x1234567ABC).cpu().numpy()
    ref = M.poisson(x.numpy(), scale.numpy(), seed=0x1234567ABC)
    assert (got != ref).sum() <= 2
    assert np.array_equal(got[1], x[1].numpy())
    big = poisson_noise(torch.full((1, 1, 256, 256), 0.5).cuda(), torch.tensor([10.0]), seed=7).cpu().numpy() * 10.0
    assert abs(big.mean() - 5.0) < 0.05 and abs(big.var() - 5.0) < 0.15


@pytest.mark.gpu
def test_degrade_batch_every_recipe():
    from mp_hsir_b200.degrade import ALL_RECIPES
    clean = synthetic_input((48, 31, 64, 64), seed=8).cuda()
    cirrus = synthetic_input((1, 1, 64, 64), seed=21)[0, 0]
    with pytest.raises(ValueError):
        degrade_batch(clean, seed=13, de_types=ALL_RECIPES, generator=torch.Generator().manual_seed(5))   # haze needs its map
    noisy, tid = degrade_batch(clean, seed=13, de_types=ALL_RECIPES, generator=torch.Generator().manual_seed(5), cirrus=cirrus)
    kinds = [ALL_RECIPES[int(t)] for t in tid.view(-1)]
    assert set(kinds) == set(ALL_RECIPES)
    for b, kind in enumerate(kinds):
        assert not torch.equal(noisy[b], clean[b]) and bool(torch.isfinite(noisy[b]).all())
        if kind == "circle_blur":
            assert float(noisy[b].var()) < float(clean[b].var())
        elif kind == "poissonN":
            c = noisy[b] * 10.0
            assert torch.equal(c.round(), c) or float((c.round() - c).abs().max()) < 1e-4


@pytest.mark.gpu
def test_haze_matches_oracle():
    """mphsir_topk_mean + mphsir_haze vs the float64 restatement of utils/degradation_utils.py:252-273: a 64 x 64 patch (top_k = 1,
    the band maximum), a 128 x 160 scene (top_k = 2) and top_percent raised so that top_k = 40 with tied values; per-sample
    cirrus maps; samples with omega 0 untouched"""
    from mp_hsir_b200 import lib
    from mp_hsir_b200.degrade import haze
    for (B, C, H, W), pct in (((3, 31, 64, 64), 0.01), ((2, 100, 128, 160), 0.01), ((2, 5, 64, 64), 1.0)):
        clean = synthetic_input((B, C, H, W), seed=12)
        if pct == 1.0:
            clean = (clean * 50).round() / 50                  # many ties among the brightest pixels
        cirrus = synthetic_input((B, 1, H, W), seed=13)[:, 0] * 1.6     # some pixels with 1 - omega * cirrus <= 0
        omega = torch.tensor([0.75, 0.0, 1.0][:B])
        k = max(int(H * W * pct / 100), 1)
        light = lib.topk_mean(clean.cuda().contiguous(), k).cpu().numpy()
        out = torch.full_like(clean, 7.0).cuda()
        haze(clean.cuda(), cirrus, omega, top_percent=pct, out=out)
        out = out.cpu().numpy()
        for b in range(B):
            flat = clean[b].numpy().reshape(C, -1)
            want = np.sort(flat, axis=1)[:, -k:].astype(np.float64).mean(axis=1)
            assert abs(light[b] - want).max() < 1e-6
            if float(omega[b]) == 0:
                assert (out[b] == 7.0).all()
            else:
                ref = M.haze(clean[b].numpy(), cirrus[b].numpy(), float(omega[b]), top_percent=pct)
                assert abs(out[b] - ref).max() < 2e-6


@pytest.mark.gpu
def test_degrade_structured_matches_oracle_stream():
    from mp_hsir_b200.degrade import degrade_structured, draw_structured
    B, C, H, W = 6, 9, 24, 40
    x = synthetic_input((B, C, H, W), seed=3)
    code = torch.tensor([1, 1, 0, 1, 1, 1])
    colmul, coladd, impulse, active = draw_structured(code, C, W, torch.Generator().manual_seed(8))
    got = degrade_structured(x.clone().cuda(), colmul, coladd, impulse, active, seed=0xABCDEF0123).cpu().numpy()
    ref = M.degrade_structured(x.numpy(), colmul.numpy(), coladd.numpy(), impulse.numpy(), active.numpy(), seed=0xABCDEF0123)
    assert np.array_equal(got, ref)          # integer decisions + one fp32 multiply-add: bit exact
    assert np.array_equal(got[2], x[2].numpy())


@pytest.mark.gpu
def test_degrade_batch_complex_full():
    clean = synthetic_input((24, 31, 64, 64), seed=6).cuda()
    noisy, tid = degrade_batch(clean, seed=5, generator=torch.Generator().manual_seed(2), complex_full=True)
    plain, tid2 = degrade_batch(clean, seed=5, generator=torch.Generator().manual_seed(2))
    assert torch.equal(tid, tid2)
    for b in range(24):
        if int(tid[b, 0]) == 1:
            assert not torch.equal(noisy[b], plain[b])       # the structured half changed something
        else:
            assert torch.equal(noisy[b], plain[b])
