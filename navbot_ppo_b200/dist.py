"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests) for the two exchanges the path needs:

  * once per iteration: all-reduce of (sum A, sum A^2, count) so every rank normalises the
    advantages with the GLOBAL mean / unbiased std (ppo.py:284);
  * once per epoch: all-reduce (sum) of the flat gradient, 100,608 fp32 = 402 KB.  Each rank's
    gradient already carries the 1/n_global factor of the reference's .mean() losses
    (ppo.py:342-343), so the sum IS the full-batch gradient and every rank applies the same
    Adam step to identical weights.

Agents shard by contiguous global id ranges; the simulator's RNG is keyed by global id, so a
sharded run reproduces the single-GPU episode stream agent for agent.
"""
from __future__ import annotations

import torch
import torch.distributed as td


def is_dist() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_dist() else 1


def rank() -> int:
    return td.get_rank() if is_dist() else 0


def shard_range(n_total: int, rank_: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the global agent ids owned by `rank_` (contiguous, sizes differ by <= 1)."""
    base, rem = divmod(n_total, world)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist() and td.get_world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if is_dist() and td.get_world_size() > 1:
        td.broadcast(t, src=src)
    return t


def advantage_moments(local_sum: float, local_sq: float, local_n: float, device=None) -> tuple[float, float, float]:
    """Global (mean, unbiased std, n) from per-rank (sum, sum of squares, count)."""
    s = torch.tensor([local_sum, local_sq, local_n], dtype=torch.float64, device=device)
    all_reduce_sum_(s)
    tot, sq, n = (float(x) for x in s)
    mean = tot / n
    var = max((sq - n * mean * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5, n
