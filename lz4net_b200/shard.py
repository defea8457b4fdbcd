"""Multi-GPU sharding of a batch of independent blocks.

lz4net's blocks are independent by construction (doc/compatibility.md:4-7), so the multi-GPU form of the path is plain
data parallelism: rank r owns the contiguous block range [r*B, (r+1)*B) (weak scaling: B blocks per GPU) or an even
split of a fixed total (strong scaling, used by LZ4Stream-style callers to keep stream order on gather).  There is no
data-path collective; torch.distributed is only used for the barrier and for the max-over-ranks of the timings.
"""
from __future__ import annotations

from typing import Sequence, Tuple


def weak_range(rank: int, blocks_per_rank: int) -> Tuple[int, int]:
    """Global block indices owned by `rank` when every rank processes `blocks_per_rank` blocks."""
    return rank * blocks_per_rank, (rank + 1) * blocks_per_rank


def strong_range(rank: int, world: int, total_blocks: int) -> Tuple[int, int]:
    """Contiguous, order-preserving split of `total_blocks` over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(total_blocks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(values: Sequence[float], device=None):
    """Max over ranks of each value (device timings: the slowest rank defines the job's time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def reduce_sum(values: Sequence[float], device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(bytes_per_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job bytes/s: all ranks' bytes over the slowest rank's time."""
    total = reduce_sum([bytes_per_rank], device)[0]
    worst = reduce_max([seconds_this_rank], device)[0]
    return total / worst
