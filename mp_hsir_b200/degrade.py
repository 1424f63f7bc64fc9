"""Training-time degradations synthesised on the device.

``ImageTransformDataset.__getitem__`` (utils/dataset_utils.py:128-146) draws ONE degradation per sample in 16 NumPy
DataLoader workers per rank (options.py:21); at ~800 patches/s per GPU those workers, not the model, bound the step.  The
four array-only recipes of its ``de_dict`` (:112) are one elementwise pass here (``mphsir_degrade``): the batch travels to
the GPU clean, the per-sample parameters (a few floats) are drawn on the host exactly like the reference draws them, and
the B*C*H*W random numbers come from a counter-based Philox4x32-10 stream inside the kernel:

    gaussianN : sigma ~ U(30, 70) / 255, iid over the cube              (degradation_utils.py:25-31)
    complexN  : its non-iid Gaussian part: per band sigma from {10,30,50,70}/255   (:33-39)
    inpaint   : rand(C,H,W) > ratio, ratio from {0.7, 0.8, 0.9}          (:227-233)
    bandmiss  : int(ratio*C) whole bands zeroed, ratio from {0.1,0.2,0.3} (:275-284)

    complexN (complete, ``complex_full=True``): after the non-iid Gaussian noise ONE of deadline / impulse / stripe noise on a
                third of the bands (:296-316, :41-84) — drawn per sample on the host (band choice, column positions, amounts; a
                few KB) and applied by a second in-place launch (``mphsir_degrade_structured``)
    blur      : (optional fifth recipe) Gaussian blur, kernel size from {9, 15, 21}, sigma 0.3((k-1)/2 - 1) + 0.8, zero
                padding (:91-108) — a second launch (``mphsir_gaussian_blur``) over the blur samples only
    sr        : (optional sixth recipe) bicubic down-sampling by 2, 4 or 8 (``F.interpolate(mode='bicubic', align_corners=True)``,
                :165-176) and f x f pixel replication back to the patch size (:189-200, chained by single_degrade :431-432) —
                one launch (``mphsir_sr_degrade``) over the sr samples only.  With it the reference's DEFAULT natural-scene list
                (options.py:15: gaussianN, complexN, blur, sr, inpaint, bandmiss) is synthesised entirely on the device
    circle_blur / poissonN : (remote-sensing extras, dataset_utils.py:117) the disc-cut Gaussian blur (:110-128; host-built k x k
                kernel through the generic ``mphsir_blur2d``, which also serves square / motion blur kernels, :130-163) and Poisson
                noise Poisson(x * scale) / scale (:86-89; ``mphsir_poisson``, a third Philox stream) — ``draw_recipes``
    haze      : (remote-sensing default list, options.py:17) x T + A (1 - T) with the transmission of a cirrus-band map (:235-273);
                the map itself comes from the reference's .mat files, so the caller passes it (``cirrus=``) — on the device:
                the per-band atmospheric light (``mphsir_topk_mean``) and the elementwise pass (``mphsir_haze``)

The task id of a sample is the index of its degradation in the active ``de_type`` list, shape [B,1] (:140).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import lib

RECIPES = ("gaussianN", "complexN", "inpaint", "bandmiss")          # the default set (one elementwise launch)
ALL_RECIPES = RECIPES + ("blur", "sr", "circle_blur", "poissonN", "haze")    # + recipes that need their own kernel
REFERENCE_DEFAULT_RS = ("gaussianN", "complexN", "blur", "sr", "inpaint", "haze", "bandmiss")   # options.py:17 (remote sensing)
REFERENCE_DEFAULT = ("gaussianN", "complexN", "blur", "sr", "inpaint", "bandmiss")   # options.py:15 (task ids 0..5 in this order)
DE_RANGE = {"gaussianN": (30.0, 70.0), "complexN": (10.0, 30.0, 50.0, 70.0), "inpaint": (0.7, 0.8, 0.9), "bandmiss": (0.1, 0.2, 0.3),
            "blur": (9, 15, 21), "sr": (2, 4, 8), "circle_blur": (9,), "poissonN": (10.0,), "haze": (0.5, 0.75, 1.0)}


def draw_parameters(B: int, C: int, de_types: Sequence[str] = RECIPES, generator: Optional[torch.Generator] = None,
                    with_blur: bool = False, with_sr: bool = False):
    """Host-side draws of one degradation per sample -> (task_id [B,1] int64, sigma [B,C], keep [B,C], mask_ratio [B]);
    a few hundred bytes and a dozen vectorised torch calls, the only per-step host work.  with_blur: one more element, the
    Gaussian-blur kernel size per sample (int32 [B], 0 = not a blur sample); with_sr: one more (after it), the down-sampling
    factor per sample (int32 [B], 0 = not an sr sample)."""
    g = generator
    for k in de_types:
        if k not in ALL_RECIPES:
            raise ValueError(f"{k!r} is not an array-only recipe ({ALL_RECIPES})")
    if "blur" in de_types and not with_blur:
        raise ValueError("the 'blur' recipe needs with_blur=True (its kernel sizes are a fifth return value)")
    if "sr" in de_types and not with_sr:
        raise ValueError("the 'sr' recipe needs with_sr=True (its down-sampling factors are one more return value)")
    tid = torch.randint(0, len(de_types), (B, 1), generator=g)
    code = torch.tensor([ALL_RECIPES.index(k) for k in de_types])[tid[:, 0]]      # recipe of every sample
    r = DE_RANGE
    # gaussianN: sigma = U(30, 70) / 255 for the whole cube (degradation_utils.py:26-27)
    sig_g = (r["gaussianN"][0] + (r["gaussianN"][1] - r["gaussianN"][0]) * torch.rand(B, generator=g)) / 255.0
    # complexN (non-iid part): one sigma per band from the list (:34-35)
    tab = torch.tensor(r["complexN"])
    sig_c = tab[torch.randint(0, len(tab), (B, C), generator=g)] / 255.0
    sigma = torch.where((code == 0)[:, None], sig_g[:, None].expand(B, C), torch.where((code == 1)[:, None], sig_c, torch.zeros(B, C)))
    # inpaint: keep a pixel when rand > ratio (:230); -1 = no mask (u > -1 always)
    rt = torch.tensor(r["inpaint"])[torch.randint(0, len(r["inpaint"]), (B,), generator=g)]
    ratio = torch.where(code == 2, rt, torch.full((B,), -1.0))
    # bandmiss: int(ratio * C) bands chosen without replacement (:277-279)
    pct = torch.tensor(r["bandmiss"], dtype=torch.float64)[torch.randint(0, len(r["bandmiss"]), (B,), generator=g)]
    n_lost = (pct * C).floor().to(torch.int64)
    order = torch.rand(B, C, generator=g).argsort(dim=1).argsort(dim=1)          # a random permutation rank per band
    keep = ((order >= n_lost[:, None]) | (code != 3)[:, None]).to(torch.float32)
    ret = [tid, sigma.contiguous(), keep.contiguous(), ratio.contiguous()]
    if with_blur:
        # blur: kernel size from the list (dataset_utils.py:112 'blur': [(9, 15, 21)]); blur samples pass the elementwise kernel unchanged
        ks = torch.tensor(r["blur"], dtype=torch.int32)[torch.randint(0, len(r["blur"]), (B,), generator=g)]
        ret.append(torch.where(code == 4, ks, torch.zeros(B, dtype=torch.int32)).contiguous())
    if with_sr:
        # sr: intensity = randint(0, 2) picks the factor from (2, 4, 8) (degradation_utils.py:373-374); sr samples pass the
        # elementwise kernel unchanged as well
        fs = torch.tensor(r["sr"], dtype=torch.int32)[torch.randint(0, len(r["sr"]), (B,), generator=g)]
        ret.append(torch.where(code == 5, fs, torch.zeros(B, dtype=torch.int32)).contiguous())
    return tuple(ret)


def draw_recipes(B: int, C: int, de_types: Sequence[str], generator: Optional[torch.Generator] = None) -> dict:
    """``draw_parameters`` for any subset of ALL_RECIPES, as a dict: tid, sigma, keep, ratio and — for the recipes present in
    `de_types` — ksize (blur), factor (sr), circle (kernel size per sample, 0 = not a circle-blur sample), poisson (scale per
    sample, 0 = not a Poisson sample), omega (haze strength per sample, 0 = not a haze sample)."""
    wb, wsr = "blur" in de_types, "sr" in de_types
    tid, sigma, keep, ratio, *extra = draw_parameters(B, C, de_types, generator, with_blur=wb, with_sr=wsr)
    d = {"tid": tid, "sigma": sigma, "keep": keep, "ratio": ratio}
    if wb:
        d["ksize"] = extra.pop(0)
    if wsr:
        d["factor"] = extra.pop(0)
    code = torch.tensor([ALL_RECIPES.index(k) for k in de_types])[tid[:, 0]]
    if "circle_blur" in de_types:
        ks = torch.tensor(DE_RANGE["circle_blur"], dtype=torch.int32)[torch.randint(0, len(DE_RANGE["circle_blur"]), (B,), generator=generator)]
        d["circle"] = torch.where(code == ALL_RECIPES.index("circle_blur"), ks, torch.zeros(B, dtype=torch.int32)).contiguous()
    if "poissonN" in de_types:
        sc = torch.tensor(DE_RANGE["poissonN"])[torch.randint(0, len(DE_RANGE["poissonN"]), (B,), generator=generator)]
        d["poisson"] = torch.where(code == ALL_RECIPES.index("poissonN"), sc, torch.zeros(B)).contiguous()
    if "haze" in de_types:
        om = torch.tensor(DE_RANGE["haze"])[torch.randint(0, len(DE_RANGE["haze"]), (B,), generator=generator)]
        d["omega"] = torch.where(code == ALL_RECIPES.index("haze"), om, torch.zeros(B)).contiguous()
    return d


def circle_kernel(kernel_size: int) -> torch.Tensor:
    """the k x k kernel of `_apply_circle_blur` (utils/degradation_utils.py:111-120): exp(-d^2 / (2 r^2)) inside the disc of radius
    r = k // 2 around the centre, 0 outside, normalised; fp32 [k, k] on the host"""
    k = int(kernel_size)
    r = k // 2
    ax = torch.arange(k, dtype=torch.float64) - r
    d2 = ax[:, None] ** 2 + ax[None, :] ** 2
    kern = torch.where(d2.sqrt() <= r, torch.exp(-d2 / (2.0 * r * r)), torch.zeros(())).to(torch.float32)
    return kern / kern.sum()


def blur2d(clean: torch.Tensor, kernel: torch.Tensor, active: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depthwise k x k blur of the samples with active[b] != 0 (all when None) with ONE kernel, zero padding k // 2 — circle blur
    (`circle_kernel`), square blur (`torch.full((k, k), 1 / k**2)`) or a motion-blur kernel built with cv2 as the reference does
    (utils/degradation_utils.py:110-163).  Other samples of `out` keep their content (a fresh `out` starts as a copy of `clean`)."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    k = kernel.shape[-1]
    if kernel.dim() != 2 or kernel.shape[0] != k or k % 2 == 0 or k > 21:
        raise ValueError(f"blur2d: a square kernel of odd size <= 21 is needed, got {tuple(kernel.shape)}")
    c = clean.detach().float().contiguous()
    if out is None:
        out = c.clone()
    act = torch.ones(c.shape[0], dtype=torch.int32) if active is None else (active != 0).to(torch.int32)
    if int(act.sum()) > 0:
        with torch.cuda.device(c.device):
            lib.blur2d(c, out, kernel.to(device=c.device, dtype=torch.float32).contiguous(), act.to(c.device).contiguous())
    return out


def poisson_noise(clean: torch.Tensor, scale: torch.Tensor, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Poisson(max(x, 0) * scale[b]) / scale[b] for the samples with scale[b] > 0 (utils/degradation_utils.py:86-89); other
    samples of `out` keep their content (a fresh `out` starts as a copy of `clean`)."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    c = clean.detach().float().contiguous()
    if out is None:
        out = c.clone()
    if float(scale.max()) > 0:
        with torch.cuda.device(c.device):
            lib.poisson(c, out, scale.to(device=c.device, dtype=torch.float32).contiguous(), seed)
    return out


def haze(clean: torch.Tensor, cirrus: torch.Tensor, omega: torch.Tensor, gamma: float = 1.0, top_percent: float = 0.01,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Haze for the samples with omega[b] > 0 (utils/degradation_utils.py:252-273).  cirrus: the cirrus-band map(s) at the patch
    resolution, [H,W] (shared) or [B,H,W] — what the reference loads from its .mat files and cv2.resize's (:237-253).  The
    wavelength grid is the reference's hard-coded linspace(400, 1000, 100), so C <= 100.  Other samples of `out` keep their
    content (a fresh `out` starts as a copy of `clean`)."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    c = clean.detach().float().contiguous()
    B, C, H, W = c.shape
    if C > 100:
        raise ValueError("haze: the reference's wavelength grid has 100 entries (C <= 100)")
    if cirrus.dim() == 2:
        cirrus = cirrus[None].expand(B, H, W)
    if tuple(cirrus.shape) != (B, H, W):
        raise ValueError(f"haze: cirrus map of shape {tuple(cirrus.shape)} for a batch {tuple(c.shape)}")
    if out is None:
        out = c.clone()
    if float(omega.max()) > 0:
        wl = torch.linspace(400, 1000, 100, dtype=torch.float64)
        expo = ((wl[0] / wl[:C]) ** gamma).to(torch.float32)
        f = lambda t: t.to(device=c.device, dtype=torch.float32).contiguous()  # noqa: E731
        with torch.cuda.device(c.device):
            light = lib.topk_mean(c, max(int(H * W * top_percent / 100), 1))
            lib.haze(c, out, f(cirrus), f(omega), f(expo), light)
    return out


def draw_structured(code: torch.Tensor, C: int, W: int, generator: Optional[torch.Generator] = None):
    """The structured half of complexN for the samples with code == 1 (utils/degradation_utils.py:296-316):
    -> (colmul [B,C,W], coladd [B,C,W], impulse [B,C], active [B] int32).  Vectorised restatement of the reference's per-band
    loops: a third of the bands (floor(C/3), a random subset: `np.random.permutation(B)[:num_bands]`), per chosen band n random
    columns (`np.random.permutation(range(W))[:n]`) with n ~ randint(ceil(.05 W), ceil(.15 W)) for deadline (:61) and
    randint(floor(.05 W), floor(.15 W)) for stripes (:48), stripe offsets U(0,1) * 0.5 - 0.25 subtracted (:52-53), impulse
    amount from (0.1, 0.3, 0.5, 0.7) (:302, dataset_utils.py:112)."""
    import math
    g = generator
    B = code.numel()
    is_c = code == 1
    ctype = torch.randint(0, 3, (B,), generator=g)                                   # 0 deadline, 1 impulse, 2 stripe (:304)
    band_rank = torch.rand(B, C, generator=g).argsort(dim=1).argsort(dim=1)
    band_sel = band_rank < (C // 3)                                                  # floor(1/3 * B) bands
    col_rank = torch.rand(B, C, W, generator=g).argsort(dim=2).argsort(dim=2)        # a random permutation rank per column
    lo_d, hi_d = math.ceil(0.05 * W), math.ceil(0.15 * W)
    lo_s, hi_s = math.floor(0.05 * W), math.floor(0.15 * W)
    n_dead = torch.randint(lo_d, max(hi_d, lo_d + 1), (B, C), generator=g)
    n_stripe = torch.randint(lo_s, max(hi_s, lo_s + 1), (B, C), generator=g)
    dead = (col_rank < n_dead[:, :, None]) & band_sel[:, :, None] & (is_c & (ctype == 0))[:, None, None]
    stripe = (col_rank < n_stripe[:, :, None]) & band_sel[:, :, None] & (is_c & (ctype == 2))[:, None, None]
    offs = torch.rand(B, C, W, generator=g) * 0.5 - 0.25
    colmul = (~dead).to(torch.float32)
    coladd = torch.where(stripe, -offs, torch.zeros(()))
    amount = torch.tensor((0.1, 0.3, 0.5, 0.7))[torch.randint(0, 4, (B,), generator=g)]
    impulse = torch.where(band_sel & (is_c & (ctype == 1))[:, None], amount[:, None].expand(B, C), torch.zeros(()))
    return colmul.contiguous(), coladd.contiguous(), impulse.contiguous(), is_c.to(torch.int32).contiguous()


def degrade_structured(x: torch.Tensor, colmul: torch.Tensor, coladd: torch.Tensor, impulse: torch.Tensor, active: torch.Tensor,
                       seed: int) -> torch.Tensor:
    """in place on the device batch x (the output of `degrade`): deadline / stripe / impulse noise of the active samples"""
    if not x.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    dev = x.device
    f = lambda t: t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()  # noqa: E731
    with torch.cuda.device(dev):
        lib.degrade_structured(x, f(colmul), f(coladd), f(impulse), active.to(device=dev, dtype=torch.int32).contiguous(), seed)
    return x


def degrade(clean: torch.Tensor, sigma: torch.Tensor, keep: torch.Tensor, mask_ratio: torch.Tensor, seed: int,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """clean [B,C,H,W] fp32 on the device -> degraded batch (one libmphsir launch); parameters may live on the host."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    dev = clean.device
    c = clean.detach().float().contiguous()
    out = torch.empty_like(c) if out is None else out
    f = lambda t: t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()  # noqa: E731
    with torch.cuda.device(dev):
        lib.degrade(c, out, f(sigma), f(keep), f(mask_ratio), seed)
    return out


def gaussian_blur(clean: torch.Tensor, ksize: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gaussian-blur the samples with ksize[b] > 0 (utils/degradation_utils.py:91-108); other samples of `out` keep their content
    (a fresh `out` starts as a copy of `clean`)."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    c = clean.detach().float().contiguous()
    if out is None:
        out = c.clone()
    kmax = int(ksize.max()) if ksize.numel() else 0
    if kmax > 0:
        with torch.cuda.device(c.device):
            lib.gaussian_blur(c, out, ksize.to(device=c.device, dtype=torch.int32).contiguous(), kmax)
    return out


def sr_degrade(clean: torch.Tensor, factor: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Bicubic down-sampling by factor[b] + replication back to H x W for the samples with factor[b] > 0
    (utils/degradation_utils.py:165-176, :189-200); other samples of `out` keep their content (a fresh `out` starts as a copy
    of `clean`)."""
    if not clean.is_cuda:
        raise RuntimeError("mp_hsir_b200.degrade runs on a CUDA device via libmphsir.so; there is no CPU fallback")
    c = clean.detach().float().contiguous()
    if out is None:
        out = c.clone()
    fmax = int(factor.max()) if factor.numel() else 0
    if fmax > min(c.shape[2], c.shape[3]):
        raise ValueError(f"sr factor {fmax} exceeds the patch size {tuple(c.shape[2:])}")
    if fmax > 0:
        with torch.cuda.device(c.device):
            lib.sr_degrade(c, out, factor.to(device=c.device, dtype=torch.int32).contiguous())
    return out


def degrade_batch(clean: torch.Tensor, seed: int, de_types: Sequence[str] = RECIPES,
                  generator: Optional[torch.Generator] = None, complex_full: bool = False,
                  cirrus: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(degraded [B,C,H,W], task_id [B,1]) for a clean device batch — what the DataLoader's collate hands train.py:50-58.
    complex_full: complexN samples also get their deadline / impulse / stripe half (one more in-place launch)."""
    B, C, W = clean.shape[0], clean.shape[1], clean.shape[3]
    d = draw_recipes(B, C, de_types, generator)
    tid = d["tid"]
    # blur / sr / circle_blur / poissonN samples leave the elementwise pass as copies of the clean patch
    out = degrade(clean, d["sigma"], d["keep"], d["ratio"], seed)
    if complex_full and "complexN" in de_types:
        code = torch.tensor([ALL_RECIPES.index(k) for k in de_types])[tid[:, 0]]
        degrade_structured(out, *draw_structured(code, C, W, generator), seed=seed)
    if "ksize" in d:
        gaussian_blur(clean, d["ksize"], out=out)
    if "factor" in d:
        sr_degrade(clean, d["factor"], out=out)
    if "circle" in d:
        for k in sorted(set(d["circle"].tolist()) - {0}):
            blur2d(clean, circle_kernel(k), d["circle"] == k, out=out)
    if "poisson" in d:
        poisson_noise(clean, d["poisson"], seed, out=out)
    if "omega" in d:
        if cirrus is None:
            raise ValueError("the 'haze' recipe needs the cirrus-band map(s) of the batch (cirrus=[H,W] or [B,H,W])")
        haze(clean, cirrus, d["omega"], out=out)
    return out, tid
