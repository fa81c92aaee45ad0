"""Page-level data parallelism: pages are independent, so a job shards by page id with no data-path collective.
The only cross-rank traffic is a barrier and a max-reduction of the per-rank elapsed time (bench.py)."""
from __future__ import annotations


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` page ids for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def page_ids(n_items: int, rank: int, world: int) -> range:
    lo, hi = shard_bounds(n_items, rank, world)
    return range(lo, hi)


def weak_batch_seeds(pages_per_gpu: int, rank: int) -> range:
    """Seeds of the synthetic pages rank `rank` owns in the weak-scaling benchmark (fixed work per GPU)."""
    return range(rank * pages_per_gpu, (rank + 1) * pages_per_gpu)


def max_over_ranks(value: float, device=None) -> float:
    """Max-reduction of a scalar over the process group (identity without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
