"""Image-sharded data parallelism of the self-training step: one process per GPU, launched by torchrun.

The path shards by image: LRH, pseudo-label selection / refinement, the cross-entropy
terms and BatchNorm batch statistics are all per rank (the reference computes BatchNorm statistics per
forward call over the 8 images of one domain, which is exactly one rank's batch -- no SyncBN).  The
only exchanges are
  * ONE all-reduce of the flat fp32 gradient arena per step (the mean is applied by the SGD kernel
    through grad_scale = 1/world), and
  * one small all-reduce(sum) of the prototype class sums / counts, so that every rank applies the
    same EMA update and keeps identical prototypes (49 KB),
plus a one-off broadcast of rank 0's parameters at construction.  Backend: NCCL over NVLink on the
GPU box, gloo in the CPU tests (tests/test_ddp_gloo.py).

The gradient all-reduce is BUCKETED and OVERLAPPED with the backward pass (GradBuckets): the arena is cut at a few
module boundaries (PPM heads | layer4 | layer3 | the rest); the backward pass walks the arena from its end to its start,
and as soon as it enters a parameter below a boundary the bucket above it is complete and its all-reduce is launched
asynchronously (NCCL's own stream, ordered after the weight-gradient side stream) while the data-gradient chain of the
earlier layers keeps running.  Only the last, smallest bucket is exposed after the backward pass.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items: int, rank: int, world: int):
    """[lo, hi) of the contiguous shard of `n_items` images that `rank` owns (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_parameters(flat: torch.Tensor, src: int = 0):
    """make every rank start from rank `src`'s parameters (flat arena)"""
    _, world = world_info()
    if world > 1:
        dist.broadcast(flat, src=src)
    return flat


def allreduce_sum_(*tensors: torch.Tensor):
    """in-place sum over ranks; a no-op for a single process"""
    _, world = world_info()
    if world > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def grad_scale() -> float:
    """factor that turns the all-reduced SUM of per-rank mean losses' gradients into the global mean"""
    _, world = world_info()
    return 1.0 / world


def max_over_ranks(value: float, device="cpu") -> float:
    _, world = world_info()
    if world == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class GradBuckets:
    """Bucketed all-reduce of one flat gradient buffer, launched from inside the backward pass.

    `boundaries`: ascending element offsets that cut `grad` into buckets [0,b0) [b0,b1) ... [bk,n).  The backward pass
    reports every parameter it ENTERS (`reached(offset)`, offsets decrease as backward proceeds towards the input); a bucket
    is launched the first time an offset below its lower boundary is reported -- by construction of the boundaries (module
    boundaries of a sequential network) every gradient of that bucket has been produced, or at least enqueued on the streams
    `order_after` names, by then.  `finish()` launches whatever is left and makes the current stream wait for all of them."""

    def __init__(self, grad: torch.Tensor, boundaries, order_after=None):
        n = grad.numel()
        cuts = sorted({int(b) for b in boundaries if 0 < int(b) < n})
        self.grad = grad
        self.edges = [0] + cuts + [n]                    # bucket i = [edges[i], edges[i+1])
        self.order_after = order_after                   # callable -> (stream to launch on or None): see trainer
        self.next = len(self.edges) - 2                  # highest bucket not yet launched
        self.works = []
        self.active = False

    def begin(self):
        self.next = len(self.edges) - 2
        self.works = []
        self.active = world_info()[1] > 1

    def _launch(self, i):
        lo, hi = self.edges[i], self.edges[i + 1]
        if hi <= lo:
            return
        seg = self.grad[lo:hi]
        if self.order_after is not None and self.grad.is_cuda:
            side = self.order_after()                    # the collective must see the weight gradients of the side stream
            with torch.cuda.stream(side):
                self.works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))
        else:
            self.works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))

    def reached(self, offset: int):
        """the backward pass is about to produce the gradient of the parameter at `offset`"""
        if not self.active:
            return
        while self.next >= 1 and offset < self.edges[self.next]:
            self._launch(self.next)
            self.next -= 1

    def finish(self):
        if not self.active:
            return
        while self.next >= 0:
            self._launch(self.next)
            self.next -= 1
        for w in self.works:
            w.wait()                                     # CUDA: the current stream waits; gloo: blocks the host
        self.works = []
        self.active = False
