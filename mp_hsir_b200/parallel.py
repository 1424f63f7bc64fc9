"""Multi-GPU plumbing for the hot path: the forward shards over independent cubes / patches (one process per
GPU, no data-path collective — SURVEY.md §8e), so all that is needed is unit assignment and the
max-over-ranks reduction of device timings.  Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
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
