"""Checkpoint interchange with the reference's Lightning checkpoints (train.py:103, :107-116; test.py:575).

A reference checkpoint is ``{'state_dict': {'net.<key>': tensor, ...}, 'epoch': ..., 'global_step': ..., ...}`` written by
``ModelCheckpoint``.  ``load_reference_checkpoint`` applies the reference's own filter (train.py:111-116: keep keys that exist
with an identical shape, ``strict=False``), ``save_reference_checkpoint`` writes a file ``PromptIRModel.load_from_checkpoint``
/ the snippet in train.py can read back.  Host-side I/O only.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

PREFIX = "net."  # PromptIRModel.net (train.py:45)


def reference_state_dict(net: torch.nn.Module) -> Dict[str, torch.Tensor]:
    return {PREFIX + k: v.detach().cpu().clone() for k, v in net.state_dict().items()}


def save_reference_checkpoint(path: str, net: torch.nn.Module, epoch: int = 0, global_step: int = 0,
                              trainer=None) -> None:
    ckpt = {"state_dict": reference_state_dict(net), "epoch": epoch, "global_step": global_step,
            "pytorch-lightning_version": "2.0.0"}
    if trainer is not None:  # AdamW moments of the flat buffers, so that training can resume bit-for-bit
        ckpt["mp_hsir_b200_trainer"] = {"step_count": trainer.step_count, "lr": trainer.lr,
                                        "exp_avg": trainer.flat_m.detach().cpu(), "exp_avg_sq": trainer.flat_v.detach().cpu(),
                                        "param_offsets": dict(trainer.param_offsets)}
    torch.save(ckpt, path)


def _load_file(path: str, map_location, trust_pickle: bool):
    """Tensors-only unpickling first (a checkpoint is user-supplied data); Lightning checkpoints that carry arbitrary Python
    objects (hyper-parameter namespaces, callbacks) need full unpickling, which executes code from the file: opt-in only."""
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception as e:  # noqa: BLE001 - torch raises pickle.UnpicklingError / RuntimeError depending on the payload
        if not trust_pickle:
            raise RuntimeError(f"{path} cannot be read with weights_only=True ({type(e).__name__}: {e}); pass "
                               "trust_pickle=True only for a checkpoint from a source you trust") from e
        return torch.load(path, map_location=map_location, weights_only=False)


def load_reference_checkpoint(path_or_ckpt, net: torch.nn.Module, trainer=None, map_location="cpu",
                              trust_pickle: bool = False) -> dict:
    """Returns {'loaded': [...], 'skipped': [...], 'optimizer_restored': bool} (keys without the 'net.' prefix).
    `optimizer_restored` is False when a trainer was passed but the file holds no matching AdamW state (a plain
    reference checkpoint, or one written for another architecture: layout checked through `param_offsets`)."""
    ckpt = _load_file(path_or_ckpt, map_location, trust_pickle) if isinstance(path_or_ckpt, str) else path_or_ckpt
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    own = net.state_dict()
    loaded, skipped, filtered = [], [], {}
    for k, v in sd.items():
        kk = k[len(PREFIX):] if k.startswith(PREFIX) else k
        if kk in own and tuple(own[kk].shape) == tuple(v.shape):   # train.py:113
            filtered[kk] = v
            loaded.append(kk)
        else:
            skipped.append(kk)
    with torch.no_grad():  # copy in place: the trainer's parameters are views of one flat buffer and must stay so
        for kk, v in filtered.items():
            own[kk].copy_(v.to(device=own[kk].device, dtype=own[kk].dtype))
    if hasattr(net, "invalidate_packed_weights"):
        net.invalidate_packed_weights()
    st: Optional[dict] = ckpt.get("mp_hsir_b200_trainer") if isinstance(ckpt, dict) else None
    restored = False
    if trainer is not None and st is not None:
        same_layout = (st["exp_avg"].numel() == trainer.flat_m.numel()
                       and dict(st.get("param_offsets", {})) == dict(trainer.param_offsets))
        if same_layout:
            trainer.flat_m.copy_(st["exp_avg"])
            trainer.flat_v.copy_(st["exp_avg_sq"])
            trainer.step_count, trainer.lr = int(st["step_count"]), float(st["lr"])
            restored = True
        else:
            import warnings
            warnings.warn("checkpoint holds AdamW state for another parameter layout: moments and step count NOT restored")
    return {"loaded": loaded, "skipped": skipped, "optimizer_restored": restored}
