"""Image-index sharding across the GPUs of one box: the path is embarrassingly
parallel, so ranks own disjoint image ranges and never exchange pixels.  The only
communication is a gather of per-rank timing records."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of item indices owned by `rank`; blocks differ by at most 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_bytes(sizes: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks balanced by byte count (mixed-resolution test sets): block r ends at the
    first index where the running byte total reaches (r+1)/world of the grand total."""
    total = sum(sizes)
    cuts, acc, r = [0], 0, 1
    for i, s in enumerate(sizes):
        acc += s
        while r < world and acc * world >= r * total:
            cuts.append(i + 1)
            r += 1
    while len(cuts) < world:
        cuts.append(len(sizes))
    cuts.append(len(sizes))
    return [(cuts[r], max(cuts[r], cuts[r + 1])) for r in range(world)]


def gather_records(record: dict, group=None) -> List[dict]:
    """all_gather_object of one small dict per rank (timings); identity when not distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [record]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, record, group=group)
    return out
