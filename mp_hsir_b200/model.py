"""Drop-in ``MP_HSIR_Net``: same constructor, ``forward(inp_img, task_id)`` and 658-key
``state_dict`` as the reference (net/MP_HSIR.py:763-844), compute in libmphsir.so (sm_100a).

The sub-modules below are *parameter containers*: they exist so that parameter names, shapes,
registration order and default initialisation match the reference (checkpoints load with
``load_state_dict(strict=True)``); none of their ``forward`` methods is ever called.  The forward
pass is ``engine.Engine``: packed weights + a static sequence of C-ABI kernel launches on
token-major activations.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from .config import CLIP_DIM, PROMPT_LEN, SHIFT, WINDOW, NetConfig, Stage
from .synth import synthetic_clip_prompt


def _relative_position_index() -> torch.Tensor:
    """int64 [64,64] buffer of Spatial_Attention (net/MP_HSIR.py:172-182): (yp-yq+7)*15 + (xp-xq+7)."""
    t = torch.arange(WINDOW * WINDOW)
    y, x = t // WINDOW, t % WINDOW
    return (y[:, None] - y[None, :] + WINDOW - 1) * (2 * WINDOW - 1) + (x[:, None] - x[None, :] + WINDOW - 1)


def _shift_mask(res: int) -> torch.Tensor:
    """fp32 [nW,64,64] buffer of shifted PGSSTBs at the construction resolution (net/MP_HSIR.py:639-660)."""
    c = torch.arange(res)
    r = (c >= res - WINDOW).long() + (c >= res - SHIFT).long()
    lab = (3 * r[:, None] + r[None, :]).float()
    n = res // WINDOW
    lw = lab.view(n, WINDOW, n, WINDOW).permute(0, 2, 1, 3).reshape(n * n, WINDOW * WINDOW)
    d = lw[:, None, :] - lw[:, :, None]
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


class _SpatialAttnParams(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WINDOW - 1) ** 2, heads))
        self.register_buffer("relative_position_index", _relative_position_index())
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _SpectralAttnParams(nn.Module):
    """Spectral_Attention / MDTA Attention parameters (net/MP_HSIR.py:86-93, 395-402)."""

    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.temperature = nn.Parameter(torch.ones(heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=False)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, padding=1, groups=dim * 3, bias=False)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=False)


class _LocalSpectralParams(nn.Module):
    """PG_Spectral_Attention parameters (net/MP_HSIR.py:117-129)."""

    def __init__(self, dim: int, compress: int):
        super().__init__()
        r = dim // compress
        self.linear_down = nn.Linear(dim, r, bias=False)
        self.linear_up = nn.Linear(r, dim, bias=False)
        self.linear_prompt = nn.Linear(dim, PROMPT_LEN, bias=False)
        self.prompt_param = nn.Parameter(torch.rand(1, 1, PROMPT_LEN, r))
        self.q = nn.Linear(r, r, bias=False)
        self.kv = nn.Linear(r, 2 * r, bias=False)
        self.proj = nn.Linear(r, r)


class _GatedMlpParams(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden * 2)
        self.fc2 = nn.Linear(hidden, dim)


class _PGSSTBParams(nn.Module):
    """PGSSTB parameters / buffers (net/MP_HSIR.py:603-636)."""

    def __init__(self, dim: int, heads: int, compress: int, hidden: int, shift: int, construct_res: int):
        super().__init__()
        self.shift_size = shift
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _GatedMlpParams(dim, hidden)
        self.attn = _SpatialAttnParams(dim, heads)
        self.register_buffer("attn_mask", _shift_mask(construct_res) if shift > 0 else None)
        self.gobal_spectral_attn = _SpectralAttnParams(dim, heads)  # sic, reference spelling (:635)
        self.local_spectral_attn = _LocalSpectralParams(dim, compress)


class _StageParams(nn.Module):
    def __init__(self, st: Stage, hidden: int):
        super().__init__()
        self.blocks = nn.ModuleList(
            _PGSSTBParams(st.dim, st.heads, st.compress, hidden, SHIFT if i % 2 else 0, st.construct_res)
            for i in range(st.depth))


class _WithBiasLN(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.body = nn.Module()
        self.body.weight = nn.Parameter(torch.ones(dim))
        self.body.bias = nn.Parameter(torch.zeros(dim))


class _GDFNParams(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.project_in = nn.Conv2d(dim, hidden * 2, 1, bias=False)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, 3, padding=1, groups=hidden * 2, bias=False)
        self.project_out = nn.Conv2d(hidden, dim, 1, bias=False)


class _CrossAttnParams(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.temperature = nn.Parameter(torch.ones(heads, 1, 1))
        self.kv = nn.Conv2d(dim, dim * 2, 1, bias=False)
        self.kv_dwconv = nn.Conv2d(dim * 2, dim * 2, 3, padding=1, groups=dim * 2, bias=False)
        self.q = nn.Conv2d(dim, dim, 1, bias=False)
        self.q_dwconv = nn.Conv2d(dim, dim, 3, padding=1, groups=dim, bias=False)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=False)


class _CrossTransformerParams(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.norm11 = _WithBiasLN(dim)
        self.norm12 = _WithBiasLN(dim)
        self.attn = _CrossAttnParams(dim, 2)
        self.norm2 = _WithBiasLN(dim)
        self.ffn = _GDFNParams(dim, hidden)


class _TVSPParams(nn.Module):
    """TVSP parameters (net/MP_HSIR.py:539-568); text_linear / clip_linear are dead but present."""

    def __init__(self, task_classes: int, prompt_size: int, dim: int, hidden: int):
        super().__init__()
        self.prompt_size = prompt_size
        self.text_linear = nn.Linear(CLIP_DIM, dim)
        self.visual_prompt = nn.Parameter(torch.randn(1, dim, prompt_size, prompt_size))
        self.clip_linear = nn.Linear(CLIP_DIM, dim)
        self.text_prompt_learnable = nn.Parameter(torch.randn(1, task_classes, dim, 1, 1))
        self.cross_transformer = _CrossTransformerParams(dim, hidden)
        self.conv_last = nn.Conv2d(dim, dim, 3, padding=1, bias=False)


class _TransformerBlockParams(nn.Module):
    def __init__(self, dim: int, heads: int, hidden: int):
        super().__init__()
        self.norm1 = _WithBiasLN(dim)
        self.attn = _SpectralAttnParams(dim, heads)
        self.norm2 = _WithBiasLN(dim)
        self.ffn = _GDFNParams(dim, hidden)


class _PromptFusionParams(nn.Module):
    def __init__(self, dim: int, out_dim: int, heads: int, hidden: int):
        super().__init__()
        self.heads = heads
        self.transformer = _TransformerBlockParams(dim, heads, hidden)
        self.conv = nn.Conv2d(dim, out_dim, 1, bias=False)


class _Conv3x3Body(nn.Module):
    """Downsample / Upsample: `body.0` is the conv, `body.1` the parameter-free pixel (un)shuffle."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.Identity())


class _PatchEmbed(nn.Module):
    def __init__(self, cin: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, 3, padding=1, bias=False)


class _TextPrompt(nn.Module):
    """Text_Prompt (net/MP_HSIR.py:481-535): holds the constant [T,512] text embedding as a plain
    attribute (NOT in the state_dict, like the reference)."""

    def __init__(self, task_classes: int, clip_prompt: Optional[torch.Tensor]):
        super().__init__()
        self.task_classes = task_classes
        if clip_prompt is None:
            # no network for CLIP ViT-B/32 weights: cached synthetic text embeddings (north_star item 3)
            clip_prompt = synthetic_clip_prompt(task_classes)
        clip_prompt = torch.as_tensor(clip_prompt, dtype=torch.float32)
        if tuple(clip_prompt.shape) != (task_classes, CLIP_DIM):
            raise ValueError(f"clip_prompt must be [{task_classes},{CLIP_DIM}], got {tuple(clip_prompt.shape)}")
        self.clip_prompt = clip_prompt.detach().clone()

    def get_clip_prompt(self):
        return self.clip_prompt


class MP_HSIR_Net(nn.Module):
    """B200-native MP-HSIR network.  API == reference ``MP_HSIR_Net`` (net/MP_HSIR.py:763-844) plus one
    optional keyword, ``clip_prompt`` ([task_classes,512] text embeddings; defaults to the cached
    synthetic tensor because CLIP weights cannot be downloaded here) and ``precision``:
      "fp32"       tcgen05 tensor cores with operands split into bf16 hi+lo (3 MMAs, fp32 accumulate):
                   fp32-grade products, meets the 1e-4 parity bound (default)
      "fp32_exact" FFMA everywhere (bit-for-bit fp32 products; slow; numerical baseline)
      "bf16"       tcgen05 with bf16-rounded operands, fp32 accumulate/activations (1e-2 bound)"""

    def __init__(self, in_channel: int = 31, out_channel: int = 31, dim: int = 64,
                 num_blocks: Sequence[int] = (2, 4, 6), window_size: Sequence[int] = (8, 8, 8),
                 task_classes: int = 6, num_refinement_blocks: int = 4, heads: Sequence[int] = (2, 4, 8),
                 ffn_expansion_factor: float = 2.66, bias: bool = False,
                 clip_prompt: Optional[torch.Tensor] = None, precision: str = "fp32"):
        super().__init__()
        if precision not in ("fp32", "fp32_exact", "bf16"):
            raise ValueError("precision must be 'fp32' (tcgen05, bf16x3 split operands), 'fp32_exact' (FFMA) or 'bf16'")
        self.precision = precision
        cfg = NetConfig(in_channel, out_channel, dim, tuple(num_blocks), tuple(window_size), task_classes,
                        num_refinement_blocks, tuple(heads), ffn_expansion_factor, bias)
        self.cfg = cfg
        st = {s.name: s for s in cfg.stages()}
        hid = cfg.hidden

        self.patch_embed = _PatchEmbed(in_channel, dim)
        self.text_prompt = _TextPrompt(task_classes, clip_prompt)
        self.clip_prompts = self.text_prompt.get_clip_prompt()
        self.prompt1 = _TVSPParams(task_classes, 64, dim, hid(dim))
        self.prompt2 = _TVSPParams(task_classes, 32, dim * 2, hid(dim * 2))
        self.fusion1 = _PromptFusionParams(dim * 2, dim, 4, hid(dim * 2))
        self.fusion2 = _PromptFusionParams(dim * 4, dim * 2, 8, hid(dim * 4))

        self.encoder_level1 = _StageParams(st["encoder_level1"], hid(dim))
        self.down1_2 = _Conv3x3Body(dim, dim // 2)
        self.encoder_level2 = _StageParams(st["encoder_level2"], hid(dim * 2))
        self.down2_3 = _Conv3x3Body(dim * 2, dim)
        self.latent = _StageParams(st["latent"], hid(dim * 4))
        self.up3_2 = _Conv3x3Body(dim * 4, dim * 8)
        self.reduce_chan_level2 = nn.Conv2d(dim * 4, dim * 2, 1, bias=False)
        self.decoder_level2 = _StageParams(st["decoder_level2"], hid(dim * 2))
        self.up2_1 = _Conv3x3Body(dim * 2, dim * 4)
        self.decoder_level1 = _StageParams(st["decoder_level1"], hid(dim * 2))
        self.refinement = _StageParams(st["refinement"], hid(dim * 2))
        self.output = nn.Conv2d(dim * 2, out_channel, 3, padding=1, bias=False)
        self.prompts = None

        self._engine = None
        self.use_cuda_graph = False

    # -- engine management ---------------------------------------------------------------------
    def engine(self):
        from .engine import Engine
        if self._engine is None:
            self._engine = Engine(self)
        return self._engine

    def trainer(self, **optimizer_kwargs):
        """Training executor (mp_hsir_b200/train_engine.py): forward + clamp/L1 loss + hand-written backward + AdamW,
        i.e. PromptIRModel.training_step / configure_optimizers of the reference (train.py:50-69).  The parameters are
        re-homed into one flat buffer; ``param.grad`` are views of the flat gradient buffer (until the module is first called
        under autograd — ``net(x, t)`` in train mode — which hands ``param.grad`` back to torch, see autograd.py)."""
        from .train_engine import TrainEngine
        if not isinstance(self._engine, TrainEngine):
            self._engine = TrainEngine(self, **optimizer_kwargs)
        return self._engine

    def set_precision(self, precision: str) -> "MP_HSIR_Net":
        if precision not in ("fp32", "fp32_exact", "bf16"):
            raise ValueError(precision)
        self.precision = precision
        self._engine = None
        return self

    def invalidate_packed_weights(self) -> None:
        """Call after mutating parameters in place if automatic version tracking cannot see it."""
        if self._engine is not None:
            self._engine.invalidate()

    def _apply(self, fn, *a, **k):
        self._engine = None  # device / dtype moves invalidate packed weights and workspaces
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    # -- the hot path ----------------------------------------------------------------------------
    def forward(self, inp_img: torch.Tensor, task_id: Optional[torch.Tensor] = None) -> torch.Tensor:
        if task_id is None:
            raise ValueError("task_id is required (the reference dereferences it unconditionally, net/MP_HSIR.py:519)")
        if not inp_img.is_cuda:
            raise RuntimeError(
                "mp_hsir_b200.MP_HSIR_Net computes only on a CUDA sm_100a device via libmphsir.so; "
                "there is no CPU fallback (move the module and inputs to cuda)")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # train mode with gradients being recorded (train.py:58 -> loss.backward()): one autograd.Function over the whole network,
            # forward = the training launch sequence, backward = the hand-written backward kernels (autograd.py).
            # net.trainer().train_step(degraded, clean, task_id) is the fast path for the same step.
            from . import autograd as _ag
            return _ag.apply(self, inp_img, task_id)
        return self.engine().forward(inp_img, task_id)
