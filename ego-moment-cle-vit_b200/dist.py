"""Data-parallel plumbing for the moment-pooling path: one process per GPU, the batch sharded
per image (the path is independent per image - SURVEY.md section 8e), parameters replicated,
gradients all-reduced over NCCL/NVLink. There is no collective inside the path itself; the
reference's only multi-GPU code is one `nn.DataParallel` line (train.py:296-299).
"""
from __future__ import annotations

import contextlib
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import functional as EF


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


class _Pending:
    """Asynchronous all-reduce(s) of one flat buffer, issued in chunks."""

    def __init__(self, works, flat, grads, inplace, scale):
        self.works, self.flat, self.grads, self.inplace, self.scale = works, flat, grads, inplace, scale

    def wait(self) -> None:
        """Make the CURRENT stream wait for the collective (no host block with NCCL)."""
        for w in self.works:
            w.wait()
        self.works = []
        if self.scale != 1.0:
            self.flat.mul_(self.scale)
            self.scale = 1.0


class GradBuckets:
    """DDP-style gradient averaging, overlapped with the backward of the path.

    * Parameters are grouped (reverse registration order = the order backward produces them) into
      flat buckets of at most `bucket_bytes`; a bucket is averaged by ONE asynchronous all-reduce
      (NCCL `ReduceOp.AVG`; backends without AVG - gloo - sum and scale), or by `chunk_bytes`-sized
      pieces when that is set. A single contiguous gradient is reduced in place (no staging copy for
      the 302 MB `second_net.0.weight.grad`). One launch is the default on purpose: the persistent GEMM
      CTAs of the Newton-Schulz backward hold every SM, so a collective's CTAs are placed only at a
      kernel boundary - ten 32 MB launches each wait for their own boundary and measured 0.35 ms
      slower per step at 8 GPUs than one 302 MB launch (profiles/r02_n8_allreduce_variants.md).
    * `overlap=True` (default): a post-accumulate-grad hook launches a bucket the moment its last
      gradient exists. The fused `MomentHead` operator computes the weight gradient of its Linear
      FIRST and the whole Newton-Schulz backward after it inside one autograd node, so an ordinary
      hook would only fire when that node returns - after the chain. `functional.set_early_grad_hook`
      closes that gap: the operator hands the freshly written `dW` to `early_reduce` right after the
      kernel that produces it is enqueued, the all-reduce runs on NCCL's stream underneath the
      Newton-Schulz backward, and the operator makes the compute stream wait for it just before
      it returns `dW` to autograd (already averaged).
    * One backward per `reduce()`. A second backward while a bucket is in flight would add local
      gradients into a buffer that holds rank-averaged values (and race with the collective), so it
      raises; accumulate micro-batches under `no_sync()` and run the last one outside it.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None,
                 overlap: bool = True, chunk_bytes: Optional[int] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = int(bucket_bytes)
        self.chunk_bytes = int(chunk_bytes) if chunk_bytes else None
        self.overlap = bool(overlap)
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > self.bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, bucket in enumerate(self.buckets) for p in bucket}
        self._by_ptr: Dict[int, torch.nn.Parameter] = {p.data_ptr(): p for p in self.params}
        self._ready = [0] * len(self.buckets)
        self._inflight: Dict[int, _Pending] = {}
        self._early: Dict[int, _Pending] = {}      # id(param) -> reduction of the gradient handed over early
        self._sync = True
        self._hooks = []
        if self.overlap and self._active():
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            EF.set_early_grad_hook(self.early_reduce)

    # ------------------------------------------------------------------ helpers
    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _has_avg(self) -> bool:
        return dist.get_backend(self.group) == "nccl"

    def _all_reduce_chunks(self, flat: torch.Tensor, grads, inplace) -> _Pending:
        world = dist.get_world_size(self.group)
        avg = self._has_avg()
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        step = max(1, self.chunk_bytes // flat.element_size()) if self.chunk_bytes else max(1, flat.numel())
        works = [dist.all_reduce(flat[o:o + step], op=op, group=self.group, async_op=True)
                 for o in range(0, flat.numel(), step)]
        return _Pending(works, flat, grads, inplace, 1.0 if avg else 1.0 / world)

    @contextlib.contextmanager
    def no_sync(self):
        """Accumulate local gradients without communication (gradient accumulation): backward passes
        inside the context only add into `.grad`; the first backward outside it (followed by `reduce()`)
        averages the accumulated sum."""
        prev = self._sync
        self._sync = False
        try:
            yield
        finally:
            self._sync = prev

    # ------------------------------------------------------------------- hooks
    def early_reduce(self, weight_ptr: int, grad: torch.Tensor) -> Optional[_Pending]:
        """Called by the fused operator with the gradient of the parameter stored at `weight_ptr` as soon
        as the kernel producing it is enqueued. Returns a handle whose wait() the operator calls on the
        compute stream before it returns the gradient, or None when this reducer does not take it."""
        if not (self._sync and self.overlap and self._active()):
            return None
        p = self._by_ptr.get(weight_ptr)
        if p is None or p.grad is not None or not grad.is_contiguous():
            # unknown tensor, or local gradients already accumulated under no_sync(): the averaged dW
            # cannot be added to a local sum - leave it to the bucket path, which reduces the total
            return None
        if id(p) in self._early:
            raise RuntimeError("GradBuckets: second backward before reduce(); use no_sync() for accumulation")
        pend = self._all_reduce_chunks(grad.view(-1), [grad], True)
        self._early[id(p)] = pend
        return pend

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._sync:
            return
        i = self._bucket_of[id(p)]
        if i in self._inflight or self._ready[i] >= len(self.buckets[i]):
            raise RuntimeError(
                "GradBuckets: a gradient arrived for a bucket that is already being reduced - a second "
                "backward() ran before reduce(). Accumulate micro-batches under GradBuckets.no_sync() "
                "and run the last backward outside it.")
        self._ready[i] += 1
        if self._ready[i] == len(self.buckets[i]):
            self._launch(i)

    def _launch(self, i: int) -> None:
        bucket = [p for p in self.buckets[i] if id(p) not in self._early]     # handed over early: done
        if not bucket:
            self._inflight[i] = _Pending([], None, [], True, 1.0)
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
        inplace = len(grads) == 1 and grads[0].is_contiguous()
        flat = grads[0].view(-1) if inplace else torch.cat([g.reshape(-1) for g in grads])
        pend = self._all_reduce_chunks(flat, grads, inplace)
        pend.bucket = bucket
        self._inflight[i] = pend

    # ------------------------------------------------------------------ public
    def reduce(self) -> None:
        """Launch whatever is left, then make the current stream wait and write the averages back."""
        if not self._active() or not self._sync:
            return
        for i in range(len(self.buckets)):
            if i not in self._inflight:
                self._launch(i)
        for i in range(len(self.buckets)):
            pend = self._inflight.pop(i)
            pend.wait()
            if pend.flat is None:
                continue
            if not pend.inplace:
                off = 0
                for g in pend.grads:
                    n = g.numel()
                    g.copy_(pend.flat[off:off + n].view_as(g))
                    off += n
            for p, g in zip(pend.bucket, pend.grads):
                if p.grad is None:
                    p.grad = g
        for pend in self._early.values():
            pend.wait()          # normally already waited by the operator; idempotent
        self._early.clear()
        self._ready = [0] * len(self.buckets)

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if EF.get_early_grad_hook() == self.early_reduce:
            EF.set_early_grad_hook(None)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
