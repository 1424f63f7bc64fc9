"""Multi-GPU plumbing for the hot path.  Inference shards over independent cubes / patches (one process per GPU,
no data-path collective — SURVEY.md §8e): unit assignment + max-over-ranks reduction of device timings.  Training is
DDP batch sharding (train.py:118): ONE all-reduce of the flat gradient buffer per step, the mean folded into AdamW.
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of `n_units` independent cubes for `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's time (the job finishes when the last rank does)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_psnr(local: List[float], device=None) -> float:
    """Mean metric over every cube of every rank (test.py averages PSNR over the whole test set)."""
    s = sum_over_ranks(sum(local), device)
    n = sum_over_ranks(len(local), device)
    return s / max(n, 1.0)


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def all_reduce_gradients(flat_grad: torch.Tensor) -> float:
    """DDP gradient averaging for the trainer's flat gradient buffer: SUM over ranks in place (one collective over
    NVLink for the 617 tensors), returning the factor 1/world that AdamW applies to the summed gradient
    (``TrainEngine.train_step(..., world_size=w, all_reduce=all_reduce_gradients)``)."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / w
