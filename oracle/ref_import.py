"""Import the UNMODIFIED reference model: from /root/reference in the authoring container, else from the verbatim
copy ``oracle/build_ref.py`` placed under the git-ignored ``oracle/_ref/`` (which travels to the GPU box).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden*.py`` to generate the committed fixtures, by
``tests/test_reference_live.py`` (skipped when neither tree is present) and by ``bench.py --impl reference`` /
``cpu_baseline`` (the reference timed on the host cores).  Never on the product path.

Two imports of net/MP_HSIR.py are not installable here and are stubbed through
``sys.modules`` before the import:
  * ``timm.models.layers`` (net/MP_HSIR.py:11): DropPath / to_2tuple / trunc_normal_
  * ``clip`` (net/MP_HSIR.py:13, 512-515): load() -> object with encode_text, tokenize()
The CLIP stub returns ``mp_hsir_b200.synth.synthetic_clip_prompt`` — the same tensor the
CUDA module and the oracle use.
"""
from __future__ import annotations

import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_root() -> str:
    env = os.environ.get("MPHSIR_REFERENCE")
    if env:
        return env
    if os.path.isfile("/root/reference/net/MP_HSIR.py"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF_ROOT = _ref_root()


KEEP_QUEUE = None  # optional list of [B] DropPath multipliers consumed by the stub in forward-call order


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "net", "MP_HSIR.py"))


class _DropPath(torch.nn.Module):
    """timm DropPath semantics: per-sample Bernoulli keep, scaled by 1/keep_prob; identity in eval."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        if KEEP_QUEUE is not None:
            # gradient fixtures: consume pre-drawn multipliers (mask/keep_prob, [B]) in call order
            return x * KEEP_QUEUE.pop(0).view((x.shape[0],) + (1,) * (x.dim() - 1)).to(x.dtype)
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _install_stubs():
    from mp_hsir_b200.synth import synthetic_clip_prompt

    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    clip = types.ModuleType("clip")

    class _Clip:
        @staticmethod
        def encode_text(tokens):
            return synthetic_clip_prompt(tokens.shape[0])

    clip.load = lambda name, device="cpu": (_Clip(), None)
    clip.tokenize = lambda texts: torch.zeros(len(texts), 77, dtype=torch.long)
    sys.modules["clip"] = clip


def load_reference_class():
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from net.MP_HSIR import MP_HSIR_Net  # type: ignore
    return MP_HSIR_Net


def build_reference(cfg, seed: int = 0):
    """Reference module in eval mode, filled with the name-seeded synthetic weights."""
    from mp_hsir_b200.synth import fill_state_dict_
    import warnings
    cls = load_reference_class()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = cls(in_channel=cfg.in_channel, out_channel=cfg.out_channel, dim=cfg.dim,
                  num_blocks=list(cfg.num_blocks), window_size=list(cfg.window_size),
                  task_classes=cfg.task_classes, num_refinement_blocks=cfg.num_refinement_blocks,
                  heads=list(cfg.heads), ffn_expansion_factor=cfg.ffn_expansion_factor, bias=cfg.bias)
    fill_state_dict_(net, seed)
    return net.eval()
