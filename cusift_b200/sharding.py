"""Frame sharding for multi-GPU extraction (one process per GPU).

Frames are independent units, so the data path needs no collective: logical
frame f goes to rank f % world (cyclic), each rank extracts its shard with its own
context, and only the small per-frame keypoint counts are gathered for reporting.
All-pairs matching (BASELINE config 5) partitions the unordered pairs cyclically
by flattened pair index after an all-gather of the descriptor sets.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Logical frame indices owned by `rank` (cyclic; balanced to +-1 frame)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world))


def pair_index(i: int, j: int, n: int) -> int:
    """Flattened index of the unordered pair (i < j) among n sets, row-major."""
    if not (0 <= i < j < n):
        raise ValueError("need 0 <= i < j < n")
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def all_pairs(n: int) -> List[Tuple[int, int]]:
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def shard_pairs(n_sets: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Unordered pairs (i<j, query = i) owned by `rank`: cyclic over the flattened pair index."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return [p for k, p in enumerate(all_pairs(n_sets)) if k % world == rank]
