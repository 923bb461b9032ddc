"""Data-parallel plumbing for the moment-pooling path: one process per GPU, the batch sharded
per image (the path is independent per image - SURVEY.md section 8e), parameters replicated,
gradients all-reduced over NCCL/NVLink. There is no collective inside the path itself; the
reference's only multi-GPU code is one `nn.DataParallel` line (train.py:296-299).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `batch` images for `rank`; the first `batch % world` ranks
    take one extra image, so shards differ by at most one and cover the batch exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


class GradBuckets:
    """DDP-style gradient reduction: parameters are grouped (in reverse registration order, the
    order backward produces them) into flat buckets of at most `bucket_bytes`; each bucket is
    averaged with one all-reduce. `reduce()` launches the all-reduces asynchronously and returns
    after copying the averaged values back into `.grad`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)

    def reduce(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        works = []
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            if len(grads) == 1 and grads[0].is_contiguous():
                flat = grads[0].view(-1)           # large single tensors reduce in place
            else:
                flat = torch.cat([g.reshape(-1) for g in grads])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True),
                          flat, bucket, grads))
        for work, flat, bucket, grads in works:
            work.wait()
            flat.div_(world)
            if not (len(grads) == 1 and grads[0].is_contiguous()):
                off = 0
                for p, g in zip(bucket, grads):
                    n = g.numel()
                    g.copy_(flat[off:off + n].view_as(g))
                    off += n
            for p, g in zip(bucket, grads):
                if p.grad is None:
                    p.grad = g


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
