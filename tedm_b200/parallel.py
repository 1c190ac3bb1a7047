"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).

The path shards by independent units -- (image, timestep) pairs for inference / feature extraction,
sampling chains for the sampler -- so inference needs NO collective.  Training replicates the 36 M
parameter UNet and needs exactly one gradient all-reduce per step.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced partition [lo, hi) of n_units over world_size ranks (first n%w ranks get one more)."""
    base, rem = divmod(n_units, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule of bench.py: a multi-GPU number is the MAX over ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(t: torch.Tensor) -> torch.Tensor:
    """Sum a small tensor of validation accumulators (sums and counts) over ranks, in place.  Every rank then computes
    the SAME metrics from the whole validation set and takes the same control-flow decisions (best-model choice, early
    stop) -- a rank that returned early on its own shard's loss would leave the others blocked in the next all-reduce."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """Average .grad over ranks with flat fp32 buckets (one collective per bucket; the whole UNet is
    145 MB, i.e. 3 buckets -- sized for launch latency, not link count: NVSwitch gives every pair full
    bandwidth).  Returns the number of collectives issued."""
    rank, ws = world()
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    if ws == 1 or not grads:
        return 0
    n_coll, bucket, size = 0, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(ws)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll
