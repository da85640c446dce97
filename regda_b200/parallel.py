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
