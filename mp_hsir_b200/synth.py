"""Deterministic synthetic weights / text embeddings / inputs.

There is no network in the build or GPU environment, so neither the reference
checkpoints (README.md:27 of the reference) nor the CLIP ViT-B/32 text encoder
(net/MP_HSIR.py:512-515) can be obtained.  Everything parity-tested here runs on
weights produced by ``fill_state_dict_`` — a procedure that depends only on the
*name* and *shape* of each state_dict entry, so the unmodified reference module,
the CPU oracle and the CUDA module all receive bit-identical parameters without
shipping 58 MB of weights.

The values are deliberately non-degenerate (temperature != 1, LayerNorm affine
!= identity, relative-position table well away from 0) so that an op that is
skipped or mis-indexed shows up in the parity error.
"""
from __future__ import annotations

import zlib

import torch

CLIP_DIM = 512  # width of the CLIP ViT-B/32 text embedding (net/MP_HSIR.py:552)


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**31 - 1))
    return g


def synth_tensor(name: str, shape, seed: int = 0, gain: float = 0.7) -> torch.Tensor:
    """fp32 CPU tensor for state_dict entry ``name`` (pure function of name/shape/seed)."""
    g = _gen(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "temperature":
        return 0.5 + torch.rand(shape, generator=g)
    if leaf == "relative_position_bias_table":
        return 0.5 * torch.randn(shape, generator=g)
    if leaf == "prompt_param":
        return torch.rand(shape, generator=g)
    if leaf in ("visual_prompt", "text_prompt_learnable"):
        return torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight":
        if len(shape) == 1:  # LayerNorm scale
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return gain * torch.randn(shape, generator=g) / (fan_in ** 0.5)
    raise KeyError(f"no synthetic rule for state_dict entry {name!r}")


@torch.no_grad()
def fill_state_dict_(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every floating-point *parameter* of ``module`` in place.

    Buffers (relative_position_index, attn_mask) are left as constructed: they are
    functions of the architecture, not learnt.
    """
    for name, p in module.named_parameters():
        p.copy_(synth_tensor(name, p.shape, seed).to(p.dtype))


def synthetic_clip_prompt(task_classes: int, seed: int = 1234) -> torch.Tensor:
    """Stand-in for ``clip_model.encode_text(tokenize(prompts))`` -> [T, 512] fp32.

    CLIP text features have O(0.1-1) entries with norm ~10; a unit normal scaled by
    0.4 is in the same range.  Shared by the oracle stub and the CUDA module.
    """
    g = _gen(seed, f"clip_prompt/{task_classes}")
    return 0.4 * torch.randn(task_classes, CLIP_DIM, generator=g)


def synthetic_input(shape, seed: int = 0) -> torch.Tensor:
    """uniform[0,1) cube/patch batch, NCHW fp32."""
    g = _gen(seed, "input/" + "x".join(str(s) for s in shape))
    return torch.rand(tuple(shape), generator=g)


def synthetic_scene(bands: int, size: int, sigma: float = 30.0, seed: int = 0):
    """Smooth ICVL-like clean cube + Gaussian noise.

    clean = bicubic-upsampled rand(1,bands,32,32) clamped to [0,1];
    noisy = clean + N(0,(sigma/255)^2)  (utils/dataset_utils.py:293-298 of the reference).
    Returns (noisy, clean), both [1,bands,size,size] fp32.
    """
    g = _gen(seed, f"scene/{bands}/{size}")
    low = torch.rand(1, bands, 32, 32, generator=g)
    clean = torch.nn.functional.interpolate(low, size=(size, size), mode="bicubic",
                                            align_corners=False).clamp_(0, 1)
    noisy = clean + (sigma / 255.0) * torch.randn(clean.shape, generator=g)
    return noisy, clean
