"""PSNR / SSIM of the reference's evaluation loop on the device.

``test.py:180`` hands every restored cube to ``utils/val_utils.py:compute_psnr_ssim``: ``.cpu().numpy()``, then a Python loop
of skimage calls per band — 31 x (PSNR + 7x7-window SSIM) of a 512x512 plane, ~0.5 s per cube on the host, i.e. 30x the
restoration itself at 65 cubes/s.  Here both metrics of all bands come from ONE kernel pass over the two cubes on the GPU
(``mphsir_psnr_ssim``: 8 bytes per pixel, fp64 sums) and a 16-byte-per-band read-back.  Same call signatures and return
values as the reference functions, so ``test.py`` only swaps the import.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib


def _per_plane(recovered: torch.Tensor, clean: torch.Tensor):
    if recovered.shape != clean.shape:
        raise AssertionError("recovered and clean must have the same shape (val_utils.py:50)")
    if not (recovered.is_cuda and clean.is_cuda):
        raise RuntimeError("mp_hsir_b200.metrics computes on a CUDA device via libmphsir.so; there is no CPU fallback")
    shape = recovered.shape
    r = recovered.reshape(-1, *shape[-3:])
    c = clean.reshape(-1, *shape[-3:])
    B, C, H, W = r.shape
    with torch.cuda.device(r.device):
        sums = lib.psnr_ssim_sums(r, c).view(B, C, 2)
    mse = sums[..., 0] / float(H * W)
    psnr = 10.0 * torch.log10(1.0 / mse)                    # peak_signal_noise_ratio(x, y, data_range=1)
    ssim = sums[..., 1] / float((H - 6) * (W - 6))          # structural_similarity(x, y, data_range=1) defaults
    return psnr, ssim


def compute_psnr_ssim(recoverd: torch.Tensor, clean: torch.Tensor) -> Tuple[float, float, int]:
    """utils/val_utils.py:49-69: mean over bands, then over the batch -> (psnr, ssim, batch size)."""
    psnr, ssim = _per_plane(recoverd, clean)
    B = psnr.shape[0]
    vals = torch.stack([psnr.mean(dim=1).sum() / B, ssim.mean(dim=1).sum() / B]).tolist()   # one read-back
    return vals[0], vals[1], B


def compute_psnr_ssim2(recoverd: torch.Tensor, clean: torch.Tensor, degrad_patch: Optional[torch.Tensor] = None):
    """utils/val_utils.py:71-105 (band-completion eval): only bands whose degraded plane is all zero count; samples without
    such a band are skipped -> (psnr, ssim, number of samples counted)."""
    psnr, ssim = _per_plane(recoverd, clean)
    B, C = psnr.shape
    if degrad_patch is None:
        sel = torch.ones(B, C, dtype=torch.bool, device=psnr.device)
    else:
        with torch.cuda.device(psnr.device):
            sel = lib.plane_nonzero(degrad_patch.reshape(B, C, *degrad_patch.shape[-2:])).view(B, C) == 0
    n = sel.sum(dim=1)
    has = n > 0
    nn = n.clamp_min(1).to(torch.float64)
    p = (torch.where(sel, psnr, torch.zeros_like(psnr)).sum(dim=1) / nn)[has]
    s = (torch.where(sel, ssim, torch.zeros_like(ssim)).sum(dim=1) / nn)[has]
    count = int(has.sum())
    if count == 0:
        return 0, 0, 0
    return float(p.sum() / count), float(s.sum() / count), count
