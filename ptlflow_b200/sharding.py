"""Multi-GPU layout of the inference path: frame pairs shard, nothing is exchanged.

Every frame pair is independent: one process per GPU, replicated weights,
each rank takes a contiguous slice of the pairs.  The only collectives are measurement plumbing --
a barrier before the timed region and a MAX over ranks of the device-timed duration -- so this module
works identically over NCCL (GPUs) and gloo (CPU tests, world_size 2).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``n_items`` for ``rank``; sizes differ by at most one and
    the slices tile the range exactly (ragged counts and n_items < world_size are fine)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_process_group(backend: str) -> bool:
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1.  Returns True if a
    group is active."""
    _, _, world = env_rank_world()
    if world <= 1:
        return False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return True


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX-reduce a per-rank scalar (device-timed milliseconds)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
