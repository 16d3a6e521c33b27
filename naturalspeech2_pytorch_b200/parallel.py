"""Data-parallel plumbing: one process per GPU, independent samples per rank, one scalar collective.

`Model.forward` has no cross-sample op (norms are per token, attention per sample), so the path shards over the
batch with NO data-path collective; the only exchange is the all-reduce of the scalar loss.
The reference delegates this to HF accelerate / torch DDP (ns2.py:1723-1726, 1886).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of the batch owned by `rank`; sizes differ by at most one sample."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def global_mean_loss(local_loss: torch.Tensor, local_count: int) -> torch.Tensor:
    """Mean of a per-sample-mean loss over the GLOBAL batch: all-reduce of (loss * count, count).
    With equal shards this is the plain 4-byte-per-rank all-reduce(sum) / world of the north star."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_loss
    buf = torch.stack((local_loss.detach().float() * local_count,
                       torch.tensor(float(local_count), device=local_loss.device)))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[0] / buf[1]
