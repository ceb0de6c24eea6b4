"""Utterance-level data parallelism: contiguous shards, one process per GPU, no collective.

Each utterance's log-mel depends only on its own samples (its own max included), so the
path shards with no exchange step: rank ``r`` of ``W`` owns clips
``[r*N/W, (r+1)*N/W)`` (remainder spread over the first ranks).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open range of items owned by ``rank``; ranges tile ``[0, n_items)`` in rank order."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
