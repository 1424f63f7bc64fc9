"""CPU: the oracle restatements behind the metric / degradation kernels (oracle/metrics_oracle.py) against published
known answers and defining properties."""
import numpy as np
import pytest
import torch

from mp_hsir_b200.degrade import DE_RANGE, RECIPES, draw_parameters
from oracle import metrics_oracle as M


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10: counter 0 / key 0, all-ones, and the pi digits vector"""
    assert [hex(v) for v in M.philox4x32_10(np.array([0], dtype=np.uint64), 0)[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # second vector uses non-zero upper counter words (not reachable through this 64-bit counter API): check instead that
    # the stream is a function of (counter, seed) only and that seeds / counters decorrelate
    a = M.philox4x32_10(np.arange(1000, dtype=np.uint64), 1)
    b = M.philox4x32_10(np.arange(1000, dtype=np.uint64), 2)
    assert np.array_equal(a, M.philox4x32_10(np.arange(1000, dtype=np.uint64), 1)) and (a != b).mean() > 0.99
    assert len(np.unique(a)) > 3990


def test_ssim_psnr_defining_properties():
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 24, 31)).astype(np.float32)
    p, s = M.psnr_ssim(x, x.copy())
    assert np.all(np.isinf(p)) and np.allclose(s, 1.0)
    y = np.clip(x + 0.1, 0, 1)
    p, s = M.psnr_ssim(x, y)
    # constant offset inside the clip range: mse = mean(min(0.1, 1-x)^2), structure (covariances) almost intact
    mse = np.mean((np.clip(x.astype(np.float64) + np.float32(0.1), 0, 1) - x) ** 2, axis=(-1, -2))
    assert np.allclose(p, 10 * np.log10(1 / mse), rtol=1e-6) and np.all(s > 0.7) and np.all(s < 1.0)
    # values outside [0,1] are clipped first (val_utils.py:51-52)
    p2, _ = M.psnr_ssim(x * 3.0 - 1.0, np.clip(x * 3.0 - 1.0, 0, 1))
    assert np.all(np.isinf(p2))
    # brute-force SSIM of one plane: mean over every whole 7x7 window
    a, t = np.clip(x[0, 0].astype(np.float64), 0, 1), y[0, 0].astype(np.float64)
    vals = []
    for i in range(a.shape[0] - 6):
        for j in range(a.shape[1] - 6):
            wa, wt = a[i:i + 7, j:j + 7], t[i:i + 7, j:j + 7]
            ux, uy = wa.mean(), wt.mean()
            vx, vy, vxy = wa.var(ddof=1), wt.var(ddof=1), ((wa - ux) * (wt - uy)).sum() / 48.0
            vals.append(((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4)))
    assert abs(np.mean(vals) - s[0, 0]) < 1e-12


def test_band_miss_variant_counts_only_lost_bands():
    rng = np.random.default_rng(1)
    clean = rng.random((2, 4, 16, 16)).astype(np.float32)
    rec = np.clip(clean + 0.05 * rng.standard_normal(clean.shape), 0, 1).astype(np.float32)
    deg = clean.copy()
    deg[0, 1] = 0
    deg[0, 3] = 0                       # sample 1 lost nothing -> skipped
    p, s, n = M.compute_psnr_ssim2(rec, clean, deg)
    pp, ss = M.psnr_ssim(rec, clean)
    assert n == 1 and abs(p - pp[0, [1, 3]].mean()) < 1e-12 and abs(s - ss[0, [1, 3]].mean()) < 1e-12


def test_degrade_oracle_statistics_and_semantics():
    clean = np.full((3, 8, 32, 32), 0.5, dtype=np.float32)
    sigma = np.zeros((3, 8), dtype=np.float32)
    keep = np.ones((3, 8), dtype=np.float32)
    ratio = np.full(3, -1.0, dtype=np.float32)
    sigma[0] = 0.2                      # sample 0: Gaussian noise
    ratio[1] = 0.8                      # sample 1: random mask, 80 % dropped
    keep[2, [1, 5]] = 0                 # sample 2: two bands lost
    out, um, n = M.degrade(clean, sigma, keep, ratio, seed=42)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1.0) < 0.02 and abs(um.mean() - 0.5) < 0.02
    assert abs((out[0] - 0.5).std() - 0.2) < 0.01
    kept = out[1] != 0
    assert abs(kept.mean() - 0.2) < 0.02 and np.all(out[1][kept] == 0.5)
    assert np.all(out[2, [1, 5]] == 0) and np.all(out[2, [0, 2, 3, 4, 6, 7]] == 0.5)


def test_parameter_draws_follow_the_reference_ranges():
    g = torch.Generator().manual_seed(0)
    tid, sigma, keep, ratio = draw_parameters(64, 31, RECIPES, g)
    assert tid.shape == (64, 1) and tid.dtype == torch.int64 and set(tid.view(-1).tolist()) == {0, 1, 2, 3}
    for b in range(64):
        kind = RECIPES[int(tid[b, 0])]
        if kind == "gaussianN":
            assert 30 / 255 <= float(sigma[b, 0]) <= 70 / 255 and torch.all(sigma[b] == sigma[b, 0])
        elif kind == "complexN":
            assert set((sigma[b] * 255).round().tolist()) <= set(DE_RANGE["complexN"])
        elif kind == "inpaint":
            assert round(float(ratio[b]), 4) in DE_RANGE["inpaint"] and torch.all(sigma[b] == 0)
        else:
            lost = int((keep[b] == 0).sum())
            assert lost in {int(p * 31) for p in DE_RANGE["bandmiss"]} and torch.all(sigma[b] == 0) and ratio[b] < 0


def test_gaussian_blur_oracle_matches_the_reference_formula():
    """oracle gaussian_blur (explicit shifted sums) == the reference's own lines (utils/degradation_utils.py:91-108: F.conv2d with
    the outer-product kernel), evaluated here with the same torch fp32 operations because the module itself imports cv2 / skimage /
    matplotlib"""
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    clean = rng.random((5, 40, 36), dtype=np.float32)
    for kernel_size in (9, 15, 21, 7, 11):
        # the formula of :93-98 in torch fp32, as the reference evaluates it: g[i] = exp(-(i - c)^2 / (2 s^2)), s = 0.3 (c - 1) + 0.8
        c = (kernel_size - 1) / 2
        s_ = 0.3 * (c - 1) + 0.8
        g = torch.exp(-((torch.arange(kernel_size, dtype=torch.float32) - c) ** 2) / (2 * s_ ** 2))
        g = g / g.sum()
        k2 = torch.outer(g, g)
        inp = torch.from_numpy(clean)[None]
        ref = F.conv2d(inp, k2.expand(inp.shape[1], 1, -1, -1), padding=kernel_size // 2, groups=inp.shape[1])[0].numpy()
        got = M.gaussian_blur(clean, kernel_size)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-6
        # a blur preserves the mean of the interior and flattens: variance strictly drops
        assert got.var() < clean.var()


def test_blur_and_sr_parameter_draws():
    from mp_hsir_b200.degrade import ALL_RECIPES, REFERENCE_DEFAULT
    g = torch.Generator().manual_seed(1)
    ALL_RECIPES = ALL_RECIPES[:6]                              # the recipes draw_parameters itself parameterises
    tid, sigma, keep, ratio, ksize, factor = draw_parameters(96, 31, ALL_RECIPES, g, with_blur=True, with_sr=True)
    assert ksize.dtype == torch.int32 and factor.dtype == torch.int32 and set(tid.view(-1).tolist()) == {0, 1, 2, 3, 4, 5}
    for b in range(96):
        kind = ALL_RECIPES[int(tid[b, 0])]
        untouched = torch.all(sigma[b] == 0) and torch.all(keep[b] == 1) and ratio[b] < 0   # the elementwise pass copies it
        if kind == "blur":
            assert int(ksize[b]) in DE_RANGE["blur"] and int(factor[b]) == 0 and untouched
        elif kind == "sr":
            assert int(factor[b]) in DE_RANGE["sr"] and int(ksize[b]) == 0 and untouched
        else:
            assert int(ksize[b]) == 0 and int(factor[b]) == 0
    with pytest.raises(ValueError):
        draw_parameters(4, 31, ALL_RECIPES, g)
    with pytest.raises(ValueError):
        draw_parameters(4, 31, ALL_RECIPES, g, with_blur=True)
    # the reference's default natural-scene list (options.py:15): task id = position in THAT list (dataset_utils.py:134-140)
    tid, *_, ksize, factor = draw_parameters(64, 31, REFERENCE_DEFAULT, g, with_blur=True, with_sr=True)
    for b in range(64):
        kind = REFERENCE_DEFAULT[int(tid[b, 0])]
        assert (int(ksize[b]) > 0) == (kind == "blur") and (int(factor[b]) > 0) == (kind == "sr")


def test_sr_oracle_matches_the_reference_lines():
    """oracle sr_degrade (explicit cubic-convolution weights, float64) == the reference's own lines (utils/degradation_utils.py:
    165-176 `_bicubic_downsample` then :189-200 `_resize`, chained by single_degrade :431-432), re-enacted here with the same torch calls because
    the module itself imports cv2 / skimage / matplotlib"""
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    for (C, H, W) in ((5, 64, 64), (3, 32, 96), (2, 8, 8)):
        clean = rng.random((C, H, W), dtype=np.float32)
        for f in (2, 4, 8):
            low = F.interpolate(torch.from_numpy(clean)[None], size=(H // f, W // f), mode='bicubic', align_corners=True)[0]
            ref = low.repeat_interleave(f, dim=1).repeat_interleave(f, dim=2).numpy()      # `_resize`: every pixel f x f times
            got = M.sr_degrade(clean, f)
            assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-6
    # a constant cube stays constant (the four weights sum to one), and f x f blocks are flat
    const = M.sr_degrade(np.full((1, 16, 16), 0.37, dtype=np.float32), 4)
    assert np.abs(const - 0.37).max() < 1e-7
    got = M.sr_degrade(rng.random((1, 16, 16), dtype=np.float32), 4)
    assert np.array_equal(got[:, ::4, ::4].repeat(4, 1).repeat(4, 2), got)


def test_circle_blur_oracle_matches_the_reference_lines():
    """oracle circle_kernel / blur2d == utils/degradation_utils.py:110-128 re-enacted (the module itself imports cv2 /
    skimage / matplotlib); the host kernel the product uploads (degrade.circle_kernel) is the same to fp32 rounding"""
    import torch.nn.functional as F
    from mp_hsir_b200.degrade import circle_kernel
    rng = np.random.default_rng(2)
    clean = rng.random((4, 40, 36), dtype=np.float32)
    for kernel_size in (9, 5, 15):
        # the kernel the reference's double loop builds (:111-120): a Gaussian of sigma = radius on the disc of that radius,
        # float64 values stored into a float32 array, normalised by its float32 sum
        r = kernel_size // 2
        kernel = np.zeros((kernel_size, kernel_size), dtype=np.float32)
        for y, x in np.ndindex(kernel_size, kernel_size):
            d = float(np.sqrt((x - r) ** 2 + (y - r) ** 2))
            kernel[y, x] = np.exp(-(d * d) / (2 * r * r)) if d <= r else 0.0
        kernel /= kernel.sum()
        inp = torch.from_numpy(clean)[None]
        ref = F.conv2d(inp, torch.from_numpy(kernel).expand(inp.shape[1], 1, -1, -1), padding=r, groups=inp.shape[1])[0].numpy()
        assert np.array_equal(M.circle_kernel(kernel_size), kernel)
        assert np.abs(M.blur2d(clean, M.circle_kernel(kernel_size)) - ref).max() < 2e-6
        assert float((circle_kernel(kernel_size) - torch.from_numpy(kernel)).abs().max()) < 2e-8


def test_poisson_oracle_is_poisson():
    """the inversion restated in the oracle yields Poisson counts: mean = variance = lambda; inactive samples untouched; the
    counts are integers over the scale like np.random.poisson(x * scale) / scale (utils/degradation_utils.py:86-89)"""
    x = np.full((2, 1, 256, 256), 0.5, dtype=np.float32)
    x[1] = 0.3
    out = M.poisson(x, np.array([10.0, 0.0], dtype=np.float32), seed=1)
    counts = out[0] * 10.0
    assert np.array_equal(counts, np.round(counts)) and abs(counts.mean() - 5.0) < 0.05 and abs(counts.var() - 5.0) < 0.15
    assert np.array_equal(out[1], x[1])
    neg = M.poisson(np.full((1, 1, 8, 8), -0.2, dtype=np.float32), np.array([10.0], dtype=np.float32), seed=3)
    assert (neg == 0).all()                                    # np.clip(clean_patch, 0, None)


def test_draw_recipes_covers_every_recipe():
    from mp_hsir_b200.degrade import ALL_RECIPES, draw_recipes
    d = draw_recipes(128, 31, ALL_RECIPES, torch.Generator().manual_seed(0))
    kinds = [ALL_RECIPES[int(t)] for t in d["tid"].view(-1)]
    assert set(kinds) == set(ALL_RECIPES)
    for b, kind in enumerate(kinds):
        assert (int(d["circle"][b]) in DE_RANGE["circle_blur"]) == (kind == "circle_blur")
        assert (float(d["poisson"][b]) in DE_RANGE["poissonN"]) == (kind == "poissonN")
        assert (int(d["ksize"][b]) > 0) == (kind == "blur") and (int(d["factor"][b]) > 0) == (kind == "sr")
        assert (round(float(d["omega"][b]), 4) in DE_RANGE["haze"]) == (kind == "haze")


def test_haze_oracle_follows_the_reference_model():
    """utils/degradation_utils.py:255-273 on a given cirrus map: the atmospheric light is the mean of the top_k brightest pixels
    (np.partition in the reference, a sort here), a clear sky (cirrus 0) leaves the cube alone, full opacity (1 - omega cirrus <= 0)
    returns the atmospheric light, and longer wavelengths see less haze (the exponent (lambda_0 / lambda_c)^gamma falls with c)"""
    rng = np.random.default_rng(3)
    x = rng.random((6, 128, 128), dtype=np.float32)
    k = max(int(128 * 128 * 0.01 / 100), 1)
    assert k == 1
    clear = M.haze(x, np.zeros((128, 128)), 0.75)
    assert np.abs(clear - x).max() < 1e-7
    opaque = M.haze(x, np.full((128, 128), 2.0), 1.0)
    light = x.reshape(6, -1).max(axis=1)
    assert np.abs(opaque - light[:, None, None]).max() < 1e-4            # T = (1e-10)^e ~ 0
    half = M.haze(np.zeros((6, 128, 128), dtype=np.float32) + 0.2, np.full((128, 128), 0.5), 1.0)
    # constant cube: A = 0.2, so the result stays 0.2 whatever T is
    assert np.abs(half - 0.2).max() < 1e-7
    # top_k > 1: the light is the mean of the k brightest, via np.partition exactly as the reference takes it
    y = rng.random((3, 256, 256), dtype=np.float32)
    kk = max(int(256 * 256 * 0.01 / 100), 1)
    assert kk == 6
    opaque = M.haze(y, np.full((256, 256), 5.0), 1.0)
    want = np.array([np.mean(np.partition(y[c].flatten(), -kk)[-kk:]) for c in range(3)])
    assert np.abs(opaque[:, 0, 0] - want).max() < 1e-4


def test_structured_draws_follow_the_reference_counts():
    from mp_hsir_b200.degrade import draw_structured
    import math
    B, C, W = 48, 31, 64
    code = torch.tensor([1, 0, 1, 2] * 12)
    colmul, coladd, impulse, active = draw_structured(code, C, W, torch.Generator().manual_seed(3))
    assert torch.equal(active, (code == 1).int())
    kinds = set()
    for b in range(B):
        dead_bands = (colmul[b] == 0).any(dim=1)
        stripe_bands = (coladd[b] != 0).any(dim=1)
        imp_bands = impulse[b] > 0
        if code[b] != 1:
            assert not dead_bands.any() and not stripe_bands.any() and not imp_bands.any()
            continue
        assert int(dead_bands.any()) + int(stripe_bands.any()) + int(imp_bands.any()) == 1     # ONE of the three (:304-314)
        if dead_bands.any():
            kinds.add("deadline")
            assert int(dead_bands.sum()) <= C // 3
            n = (colmul[b][dead_bands] == 0).sum(dim=1)
            assert int(n.min()) >= math.ceil(0.05 * W) and int(n.max()) < math.ceil(0.15 * W)
        elif stripe_bands.any():
            kinds.add("stripe")
            assert int(stripe_bands.sum()) <= C // 3 and float(coladd[b].abs().max()) <= 0.25
            n = (coladd[b][stripe_bands] != 0).sum(dim=1)
            assert int(n.max()) < math.floor(0.15 * W)
        else:
            kinds.add("impulse")
            assert int(imp_bands.sum()) == C // 3 and round(float(impulse[b][imp_bands][0]), 4) in (0.1, 0.3, 0.5, 0.7)
    assert kinds == {"deadline", "stripe", "impulse"}


def test_degrade_structured_oracle_semantics():
    x = np.full((2, 6, 16, 20), 0.5, dtype=np.float32)
    colmul = np.ones((2, 6, 20), dtype=np.float32)
    coladd = np.zeros((2, 6, 20), dtype=np.float32)
    impulse = np.zeros((2, 6), dtype=np.float32)
    colmul[0, 1, [3, 7]] = 0
    coladd[0, 2, 5] = -0.2
    impulse[0, 4] = 0.3
    colmul[1, 0, :] = 0                       # sample 1 is inactive: nothing may change
    out = M.degrade_structured(x, colmul, coladd, impulse, np.array([1, 0]), seed=11)
    assert np.all(out[1] == 0.5)
    assert np.all(out[0, 1][:, [3, 7]] == 0) and np.all(np.delete(out[0, 1], [3, 7], axis=1) == 0.5)
    assert np.allclose(out[0, 2][:, 5], 0.3) and np.all(np.delete(out[0, 2], 5, axis=1) == 0.5)
    flipped = out[0, 4] != 0.5
    assert 0.15 < flipped.mean() < 0.45 and set(np.unique(out[0, 4][flipped]).tolist()) <= {0.0, 1.0}
    assert 0.2 < (out[0, 4][flipped] == 1.0).mean() < 0.8
