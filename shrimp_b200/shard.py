"""Read sharding over GPUs: the host-side logic of the N > 1 path.

The reference parallelises over chunks of the read stream (gmapper.c:331-608: `launch_scan_threads` hands
chunks of `chunk_size` reads -- mates of a pair stay together, :348,:400 -- to its worker threads and emits
the finished chunks in chunk order through a heap, :588-607).  Here one process per GPU plays the worker:
contiguous chunks are dealt round-robin to the ranks, every rank maps its chunks against its own replica
of the index, and rank 0 gathers the per-chunk results and emits them in chunk order.  There is no
collective on the data path; torch.distributed carries only the final gather and the timing reduction.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def chunk_plan(n_units: int, chunk_units: int, world: int) -> List[List[Tuple[int, int, int]]]:
    """Chunks of `chunk_units` units (a unit is a read, or a read PAIR so that mates stay together)
    dealt round-robin: plan[rank] = [(chunk_id, first_unit, end_unit), ...]."""
    if chunk_units <= 0 or world <= 0:
        raise ValueError("chunk_units and world must be positive")
    plan: List[List[Tuple[int, int, int]]] = [[] for _ in range(world)]
    cid = 0
    for first in range(0, n_units, chunk_units):
        plan[cid % world].append((cid, first, min(n_units, first + chunk_units)))
        cid += 1
    return plan


def map_sharded(n_units: int, chunk_units: int, map_chunk: Callable[[int, int], object], rank: int = 0,
                world: int = 1, group=None):
    """Every rank maps its chunks with map_chunk(first_unit, end_unit); rank 0 returns the per-chunk results
    in chunk order (what the reference's output heap emits), the other ranks return None."""
    mine = [(cid, map_chunk(a, b)) for cid, a, b in chunk_plan(n_units, chunk_units, world)[rank]]
    if world == 1:
        return [r for _, r in sorted(mine, key=lambda x: x[0])]
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    flat = [x for part in gathered for x in part]
    flat.sort(key=lambda x: x[0])
    assert [c for c, _ in flat] == list(range(len(flat))), "a chunk is missing from the gather"
    return [r for _, r in flat]


def max_over_ranks(value: float, world: int = 1, device=None, group=None) -> float:
    """The timed region of a multi-GPU run is the slowest rank's (bench.py)."""
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def units_per_rank(n_units: int, chunk_units: int, world: int) -> Sequence[int]:
    return [sum(b - a for _, a, b in p) for p in chunk_plan(n_units, chunk_units, world)]
