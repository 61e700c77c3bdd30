"""Data-parallel plumbing of the purification path: one process per GPU, contiguous batch shards, weights
replicated, no data-path collective.  The only exchange is ONE all-reduce of the int64[3] counters
{n_total, n_clean_correct, n_robust_correct} at the end of a run -- it replaces the reference's per-sample barrier and
four all_gathers (/root/reference/src/experiments/test_defense.py:126-127,245-248)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [lo, hi) of n samples owned by `rank`; the first n % world ranks get one extra sample.
    An image's EoT replicas are generated after sharding, so they always stay on one GPU (wrappers.py:20-22)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(x: torch.Tensor, rank: int, world: int) -> Tuple[torch.Tensor, int]:
    """-> (local slice of the batch, global index of its first sample).  The offset is what keys the Philox noise
    streams (`model.sample_offset`), which makes every sample's result independent of the number of GPUs."""
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi], lo


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """in-place SUM all-reduce of the int64[3] accuracy counters (24 bytes; NCCL over NVLink on GPUs, gloo on CPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def count_correct(logits: torch.Tensor, labels: torch.Tensor) -> int:
    return int((logits.argmax(dim=1) == labels).sum().item())
