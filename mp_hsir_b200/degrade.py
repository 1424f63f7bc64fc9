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

    blur      : (optional fifth recipe) Gaussian blur, kernel size from {9, 15, 21}, sigma 0.3((k-1)/2 - 1) + 0.8, zero
                padding (:91-108) — a second launch (``mphsir_gaussian_blur``) over the blur samples only

The task id of a sample is the index of its degradation in the active ``de_type`` list, shape [B,1] (:140).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import lib

RECIPES = ("gaussianN", "complexN", "inpaint", "bandmiss")          # the default set (one elementwise launch)
ALL_RECIPES = RECIPES + ("blur",)                                    # + recipes that need their own kernel
DE_RANGE = {"gaussianN": (30.0, 70.0), "complexN": (10.0, 30.0, 50.0, 70.0), "inpaint": (0.7, 0.8, 0.9), "bandmiss": (0.1, 0.2, 0.3),
            "blur": (9, 15, 21)}


def draw_parameters(B: int, C: int, de_types: Sequence[str] = RECIPES, generator: Optional[torch.Generator] = None,
                    with_blur: bool = False):
    """Host-side draws of one degradation per sample -> (task_id [B,1] int64, sigma [B,C], keep [B,C], mask_ratio [B]);
    a few hundred bytes and a dozen vectorised torch calls, the only per-step host work.  with_blur: a fifth element, the
    Gaussian-blur kernel size per sample (int32 [B], 0 = not a blur sample)."""
    g = generator
    for k in de_types:
        if k not in ALL_RECIPES:
            raise ValueError(f"{k!r} is not an array-only recipe ({ALL_RECIPES})")
    if "blur" in de_types and not with_blur:
        raise ValueError("the 'blur' recipe needs with_blur=True (its kernel sizes are a fifth return value)")
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
    if with_blur:
        # blur: kernel size from the list (dataset_utils.py:112 'blur': [(9, 15, 21)]); blur samples pass the elementwise kernel unchanged
        ks = torch.tensor(r["blur"], dtype=torch.int32)[torch.randint(0, len(r["blur"]), (B,), generator=g)]
        ksize = torch.where(code == 4, ks, torch.zeros(B, dtype=torch.int32))
        return tid, sigma.contiguous(), keep.contiguous(), ratio.contiguous(), ksize.contiguous()
    return tid, sigma.contiguous(), keep.contiguous(), ratio.contiguous()


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


def degrade_batch(clean: torch.Tensor, seed: int, de_types: Sequence[str] = RECIPES,
                  generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(degraded [B,C,H,W], task_id [B,1]) for a clean device batch — what the DataLoader's collate hands train.py:50-58."""
    if "blur" in de_types:
        tid, sigma, keep, ratio, ksize = draw_parameters(clean.shape[0], clean.shape[1], de_types, generator, with_blur=True)
        out = degrade(clean, sigma, keep, ratio, seed)          # blur samples leave this pass as copies of the clean patch
        return gaussian_blur(clean, ksize, out=out), tid
    tid, sigma, keep, ratio = draw_parameters(clean.shape[0], clean.shape[1], de_types, generator)
    return degrade(clean, sigma, keep, ratio, seed), tid
