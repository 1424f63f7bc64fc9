"""CPU restatements of what mp_hsir_b200/csrc/metrics.cu computes.  TEST INFRASTRUCTURE (imported by tests/ only).

* ``psnr_ssim`` — utils/val_utils.py:49-69 of the reference.  The arithmetic lives in scikit-image (a dependency of the
  reference, ``requirements.txt``: scikit-image 0.21; NOT installed here, so this restatement is "parity unpinned"
  against skimage itself): ``peak_signal_noise_ratio(x, y, data_range=1)`` = 10 log10(1 / mean((x-y)^2)) and
  ``structural_similarity(x, y, data_range=1)`` with its defaults — win_size 7, uniform_filter means, sample covariance
  (cov_norm = 49/48), K1 = 0.01, K2 = 0.03, and the mean taken over the filtered image cropped by (win_size-1)//2 = 3 —
  restated step by step below with scipy.ndimage.uniform_filter (the function skimage itself calls), in float64.
* ``philox4x32_10`` / ``degrade`` — the counter-based random stream (Salmon et al., "Parallel random numbers: as easy as
  1, 2, 3", SC'11; constants of Random123) and the elementwise degradation formula of the kernel.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def psnr_ssim(recovered: np.ndarray, clean: np.ndarray):
    """[B,C,H,W] -> (psnr [B,C], ssim [B,C]) per band plane, float64."""
    x = np.clip(recovered.astype(np.float64), 0, 1)
    y = np.clip(clean.astype(np.float64), 0, 1)
    B, C = x.shape[:2]
    psnr = np.zeros((B, C))
    ssim = np.zeros((B, C))
    K1, K2, win, R = 0.01, 0.03, 7, 1.0
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    pad = (win - 1) // 2
    for b in range(B):
        for c in range(C):
            a, t = x[b, c], y[b, c]
            psnr[b, c] = 10.0 * np.log10(R * R / np.mean((a - t) ** 2))
            ux, uy = uniform_filter(a, size=win), uniform_filter(t, size=win)
            uxx, uyy, uxy = uniform_filter(a * a, size=win), uniform_filter(t * t, size=win), uniform_filter(a * t, size=win)
            vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
            S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
            ssim[b, c] = S[pad:-pad, pad:-pad].mean()
    return psnr, ssim


def compute_psnr_ssim(recovered: np.ndarray, clean: np.ndarray):
    """utils/val_utils.py:49-69: mean over bands, then over the batch."""
    p, s = psnr_ssim(recovered, clean)
    return p.mean(axis=1).mean(), s.mean(axis=1).mean(), recovered.shape[0]


def compute_psnr_ssim2(recovered: np.ndarray, clean: np.ndarray, degrad: np.ndarray):
    """utils/val_utils.py:71-105: only bands whose degraded plane is all zero; samples without one are skipped."""
    p, s = psnr_ssim(recovered, clean)
    ps, ss, count = 0.0, 0.0, 0
    for b in range(recovered.shape[0]):
        sel = [c for c in range(recovered.shape[1]) if np.all(degrad[b, c] == 0)]
        if sel:
            ps += p[b, sel].mean()
            ss += s[b, sel].mean()
            count += 1
    return (ps / count, ss / count, count) if count else (0, 0, 0)


def philox4x32_10(counter: np.ndarray, seed: int, word2: int = 0) -> np.ndarray:
    """counter: uint64 array [n] -> uint32 [n, 4]; counter words (lo, hi, word2, 0), key (seed lo, seed hi)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c0 = counter & mask
    c1 = counter >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(word2))
    c3 = np.zeros_like(c0)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def degrade(clean: np.ndarray, sigma: np.ndarray, keep: np.ndarray, mask_ratio: np.ndarray, seed: int):
    """-> (degraded float64 [B,C,H,W], mask uniforms, normals): out = clean*keep*(u > ratio) + sigma*n"""
    B, C, H, W = clean.shape
    total = clean.size
    pairs = (total + 1) // 2
    w = philox4x32_10(np.arange(pairs, dtype=np.uint64), seed).astype(np.float64)
    f = lambda v: np.floor(v / 256.0)                      # noqa: E731  (w >> 8)
    u1 = (f(w[:, 1]) + 1.0) / 16777216.0
    u2 = f(w[:, 2]) / 16777216.0
    rad = np.sqrt(-2.0 * np.log(u1))
    n = np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], axis=1).reshape(-1)[:total]
    um = np.stack([f(w[:, 0]), f(w[:, 3])], axis=1).reshape(-1)[:total] / 16777216.0
    n, um = n.reshape(B, C, H, W), um.reshape(B, C, H, W)
    m = (um.astype(np.float32) > mask_ratio.astype(np.float32).reshape(B, 1, 1, 1)).astype(np.float64)
    out = clean.astype(np.float64) * keep.reshape(B, C, 1, 1) * m + sigma.reshape(B, C, 1, 1).astype(np.float64) * n
    return out, um, n


def gaussian_blur(clean: np.ndarray, kernel_size: int) -> np.ndarray:
    """utils/degradation_utils.py:91-108 (_apply_gaussian_blur), restated: depthwise conv of a [C,H,W] cube with the outer
    product of a normalised 1-D Gaussian (sigma = 0.3 ((k-1)/2 - 1) + 0.8, fp32 like the reference), zero padding k // 2.
    Written as explicit shifted sums in float64, so it does not share code with either implementation under test."""
    k = int(kernel_size)
    sigma = np.float32(0.3) * (np.float32(k - 1) * np.float32(0.5) - np.float32(1.0)) + np.float32(0.8)
    x = np.arange(k, dtype=np.float32)
    mean = np.float32(k - 1) / np.float32(2)
    k1 = np.exp(-((x - mean) ** 2) / (np.float32(2) * sigma ** 2)).astype(np.float32)
    k1 = (k1 / k1.sum(dtype=np.float32)).astype(np.float32)
    k2 = (k1[None, :] * k1[:, None]).astype(np.float64)          # kernel_2d of the reference (fp32 product)
    C, H, W = clean.shape
    r = k // 2
    pad = np.zeros((C, H + 2 * r, W + 2 * r), dtype=np.float64)
    pad[:, r:r + H, r:r + W] = clean
    out = np.zeros((C, H, W), dtype=np.float64)
    for dy in range(k):
        for dx in range(k):
            out += k2[dy, dx] * pad[:, dy:dy + H, dx:dx + W]     # cross-correlation = F.conv2d
    return out


def sr_degrade(clean: np.ndarray, factor: int) -> np.ndarray:
    """utils/degradation_utils.py:165-176 (_bicubic_downsample) followed by :189-200 (_resize), as single_degrade :431-432
    chains them for 'sr', restated on a [C,H,W] cube: torch's bicubic with align_corners=True is the cubic convolution with
    A = -0.75 on source coordinate i (H-1)/(h-1), rows / columns floor(src)-1 .. +2 clamped to the image; then every
    low-resolution pixel is repeated factor x factor.  Explicit weights in float64: shares no code with torch or the kernel."""
    f = int(factor)
    C, H, W = clean.shape
    h, w = H // f, W // f

    def taps(n_in: int, n_out: int):
        A = -0.75
        scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)   # fp32 like torch's opmath
        idx = np.zeros((n_out, 4), dtype=np.int64)
        wt = np.zeros((n_out, 4), dtype=np.float64)
        for i in range(n_out):
            real = np.float32(scale * np.float32(i))
            i0 = min(int(np.floor(real)), n_in - 1)
            t = float(min(max(np.float32(real - np.float32(i0)), 0.0), 1.0))
            c2 = lambda x: ((A * x - 5 * A) * x + 8 * A) * x - 4 * A          # noqa: E731  1 < |x| < 2
            c1 = lambda x: ((A + 2) * x - (A + 3)) * x * x + 1                # noqa: E731  |x| <= 1
            wt[i] = (c2(t + 1), c1(t), c1(1 - t), c2(2 - t))
            idx[i] = np.clip(i0 + np.arange(-1, 3), 0, n_in - 1)
        return idx, wt

    iy, wy = taps(H, h)
    ix, wx = taps(W, w)
    src = clean.astype(np.float64)
    rows = np.einsum("ia,ciaw->ciw", wy, src[:, iy, :])            # [C,h,W]
    low = np.einsum("jb,chjb->chj", wx, rows[:, :, ix])            # [C,h,w]
    return np.repeat(np.repeat(low, f, axis=1), f, axis=2)


def circle_kernel(kernel_size: int) -> np.ndarray:
    """utils/degradation_utils.py:111-120 (_apply_circle_blur): a Gaussian of sigma = radius cut to the disc, normalised (fp32)."""
    k = int(kernel_size)
    radius = center = k // 2
    yy, xx = np.mgrid[0:k, 0:k]
    dist = np.sqrt((xx - center) ** 2 + (yy - center) ** 2)
    kernel = np.where(dist <= radius, np.exp(-(dist ** 2) / (2 * (radius ** 2))), 0.0).astype(np.float32)
    return (kernel / kernel.sum(dtype=np.float32)).astype(np.float32)


def blur2d(clean: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """F.conv2d(x, kernel.repeat(C,1,1,1), padding=k//2, groups=C) of utils/degradation_utils.py:121-127 / :142-147 / :154-161 on a
    [C,H,W] cube: explicit shifted sums in float64 (cross-correlation, zero padding), odd k."""
    k = kernel.shape[0]
    C, H, W = clean.shape
    r = k // 2
    pad = np.zeros((C, H + 2 * r, W + 2 * r), dtype=np.float64)
    pad[:, r:r + H, r:r + W] = clean
    out = np.zeros((C, H, W), dtype=np.float64)
    for dy in range(k):
        for dx in range(k):
            out += float(kernel[dy, dx]) * pad[:, dy:dy + H, dx:dx + W]
    return out


def poisson(clean: np.ndarray, scale: np.ndarray, seed: int) -> np.ndarray:
    """utils/degradation_utils.py:86-89 (_apply_poisson) on [B,C,H,W] with the kernel's random stream: element e takes word e % 4
    of Philox block e // 4 (counter word 2 = 2), u = (w >> 8) 2^-24, and the count is the smallest n with u <= CDF_lambda(n),
    lambda = max(x, 0) * scale (fp32 product), the CDF summed term by term in float64.  scale[b] <= 0: sample untouched."""
    B = clean.shape[0]
    total = clean.size
    quads = (total + 3) // 4
    w = philox4x32_10(np.arange(quads, dtype=np.uint64), seed, word2=2).astype(np.float64)
    u = (np.floor(w / 256.0) / 16777216.0).reshape(-1)[:total].reshape(clean.shape)
    sc = scale.astype(np.float32).reshape((B,) + (1,) * (clean.ndim - 1))
    lam = (np.maximum(clean.astype(np.float32), np.float32(0)) * np.where(sc > 0, sc, np.float32(1))).astype(np.float64)
    p = np.exp(-lam)
    cdf = p.copy()
    n = np.zeros(clean.shape, dtype=np.int64)
    for it in range(1, 4097):
        go = u > cdf
        if not go.any():
            break
        n = np.where(go, it, n)
        p = np.where(go, p * lam / it, p)
        cdf = np.where(go, cdf + p, cdf)
    out = n.astype(np.float32) / np.where(sc > 0, sc, np.float32(1))
    return np.where(sc > 0, out, clean.astype(np.float32)).astype(np.float32)


def haze(hsi: np.ndarray, cirrus_band: np.ndarray, omega: float, gamma: float = 1.0, top_percent: float = 0.01) -> np.ndarray:
    """utils/degradation_utils.py:252-273 (_simulate_haze) on a [C,H,W] cube for a GIVEN cirrus-band map [H,W] (the reference
    picks a random .mat file and cv2.resize's it to the patch, :237-253 — file access, outside this restatement): atmospheric
    light = mean of the top_k brightest pixels per band, k = max(int(HW top_percent / 100), 1); t1 = 1 - omega cirrus (<= 0 ->
    1e-10); per band transmission t1 ** ((lambda_0 / lambda_c) ** gamma) with the hard-coded 100-point wavelength grid
    400..1000 nm; hazy = x T + A (1 - T).  float64 like the reference, cast to float32 at the end."""
    C, H, W = hsi.shape
    wavelength = np.linspace(400, 1000, 100)
    top_k = max(int(H * W * top_percent / 100), 1)
    flat = hsi.reshape(C, -1)
    light = np.array([np.mean(np.sort(flat[c])[-top_k:]) for c in range(C)], dtype=np.float64)
    t1 = 1 - omega * cirrus_band.astype(np.float64)
    t1 = np.where(t1 <= 0, 1e-10, t1)
    out = np.zeros(hsi.shape, dtype=np.float64)
    for c in range(C):
        expo = (wavelength[0] / wavelength[c]) ** gamma
        T = np.exp(expo * np.log(t1))
        out[c] = hsi[c] * T + light[c] * (1 - T)
    return out.astype(np.float32)


def degrade_structured(x: np.ndarray, colmul: np.ndarray, coladd: np.ndarray, impulse: np.ndarray, active: np.ndarray, seed: int):
    """utils/degradation_utils.py:41-84 on [B,C,H,W]: deadline columns (colmul 0), stripes (coladd), impulse flips with probability
    impulse[b,c] (salt with probability 1/2) from the Philox stream with counter word 2 = 1; inactive samples untouched."""
    B, C, H, W = x.shape
    total = x.size
    pairs = (total + 1) // 2
    w = philox4x32_10(np.arange(pairs, dtype=np.uint64), seed, word2=1).astype(np.float64)
    f = lambda v: np.floor(v / 256.0) / 16777216.0          # noqa: E731
    uf = np.stack([f(w[:, 0]), f(w[:, 2])], axis=1).reshape(-1)[:total].reshape(B, C, H, W).astype(np.float32)
    us = np.stack([f(w[:, 1]), f(w[:, 3])], axis=1).reshape(-1)[:total].reshape(B, C, H, W).astype(np.float32)
    v = x.astype(np.float32) * colmul.reshape(B, C, 1, W).astype(np.float32) + coladd.reshape(B, C, 1, W).astype(np.float32)
    p = impulse.reshape(B, C, 1, 1).astype(np.float32)
    flip = (uf < p) & (p > 0)
    v = np.where(flip, np.where(us < np.float32(0.5), np.float32(1), np.float32(0)), v)
    act = active.reshape(B, 1, 1, 1) != 0
    return np.where(act, v, x.astype(np.float32))
