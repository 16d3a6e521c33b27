"""Data-parallel plumbing: one process per GPU, independent samples per rank.

`Model.forward` has no cross-sample op (norms are per token, attention per sample), so the forward path shards over
the batch with NO data-path collective; the exchanges are the all-reduce of the scalar loss and, when
training, of the parameter gradients (`GradReducer`).  The reference delegates both to HF accelerate /
torch DDP (ns2.py:1723-1726, 1886).
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


class GradReducer:
    """Gradient all-reduce overlapped with the backward pass (what DDP's bucketed reducer does for the reference,
    ns2.py:1723-1726, 1886).

    `Model.grad_reducer = GradReducer()` makes the denoiser's backward hand every packed gradient buffer to `reduce()`
    the moment it is final (one transformer layer / wavenet stack at a time, in reverse order): the NCCL all-reduce
    (average over ranks) is issued asynchronously on the communication stream and runs over NVLink/NVSwitch while the
    remaining layers' dgrad / wgrad kernels execute; `finish()` (called before the gradients are returned to autograd)
    makes the compute stream wait for the outstanding collectives.  Buffers smaller than `coalesce_below` bytes are
    gathered and sent as one flat message at the end instead of one latency-bound collective each."""

    def __init__(self, group=None, coalesce_below: int = 1 << 20):
        self.group = group
        self.coalesce_below = coalesce_below
        self._pending = []
        self._small = []
        self._seen = set()
        self.bytes_reduced = 0

    @property
    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def reduce(self, t: torch.Tensor) -> None:
        base = t._base if t._base is not None else t
        key = (base.data_ptr(), base.numel())   # several gradients are views of one packed buffer: reduce it once
        if not self.active or key in self._seen or base.numel() == 0:
            return
        self._seen.add(key)
        if not base.is_contiguous():
            raise ValueError("GradReducer needs the packed gradient buffers to be contiguous")
        if base.numel() * base.element_size() < self.coalesce_below:
            self._small.append(base)
            return
        self.bytes_reduced += base.numel() * base.element_size()
        self._pending.append(dist.all_reduce(base, op=dist.ReduceOp.AVG if base.is_cuda else dist.ReduceOp.SUM,
                                             group=self.group, async_op=True))
        if not base.is_cuda:   # gloo has no AVG: scale after the wait (CPU tests)
            self._pending[-1] = (self._pending[-1], base)

    def reduce_all(self, grads: dict) -> None:
        for g in grads.values():
            self.reduce(g)

    def finish(self) -> None:
        if self._small:
            flat = torch.cat([b.reshape(-1) for b in self._small])
            cuda = flat.is_cuda
            dist.all_reduce(flat, op=dist.ReduceOp.AVG if cuda else dist.ReduceOp.SUM, group=self.group)
            if not cuda:
                flat /= dist.get_world_size(self.group)
            self.bytes_reduced += flat.numel() * flat.element_size()
            off = 0
            for b in self._small:
                b.copy_(flat[off:off + b.numel()].view_as(b))
                off += b.numel()
        for w in self._pending:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(dist.get_world_size(self.group))
            else:
                w.wait()
        self._pending, self._small, self._seen = [], [], set()
