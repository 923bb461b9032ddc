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
    """DDP-style gradient reduction. Parameters are grouped (in reverse registration order, the
    order backward produces them) into flat buckets of at most `bucket_bytes`; each bucket is
    averaged with one asynchronous all-reduce. With `overlap=True` (default) a
    post-accumulate-grad hook launches a bucket's all-reduce the moment its last gradient is
    written, so the 302 MB `second_net.0.weight` gradient - produced first in backward - travels
    over NVLink while the Newton-Schulz backward is still running. `reduce()` launches whatever is
    left, waits, and writes the averages back into `.grad`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None,
                 overlap: bool = True):
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
        self._bucket_of = {id(p): i for i, bucket in enumerate(self.buckets) for p in bucket}
        self._ready = [0] * len(self.buckets)
        self._inflight = {}
        self._hooks = []
        if overlap and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        i = self._bucket_of[id(p)]
        self._ready[i] += 1
        if self._ready[i] == len(self.buckets[i]):
            self._launch(i)

    def _launch(self, i: int) -> None:
        bucket = self.buckets[i]
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
        inplace = len(grads) == 1 and grads[0].is_contiguous()
        flat = grads[0].view(-1) if inplace else torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight[i] = (work, flat, grads, inplace)

    def reduce(self) -> None:
        if not self._active():
            return
        world = dist.get_world_size(self.group)
        for i in range(len(self.buckets)):
            if i not in self._inflight:
                self._launch(i)
        for i, bucket in enumerate(self.buckets):
            work, flat, grads, inplace = self._inflight.pop(i)
            work.wait()
            flat.div_(world)
            if not inplace:
                off = 0
                for g in grads:
                    n = g.numel()
                    g.copy_(flat[off:off + n].view_as(g))
                    off += n
            for p, g in zip(bucket, grads):
                if p.grad is None:
                    p.grad = g
        self._ready = [0] * len(self.buckets)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
