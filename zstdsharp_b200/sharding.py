"""Host scatter of independent frames across the GPUs of one box.

Frames share no state, so multi-GPU is a partition of the frame list: contiguous ranges balanced by byte weight,
one process (rank) per GPU, no data-path collective.  ``torch.distributed`` is only used by bench.py for the
barrier and the max-over-ranks timing.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(weights: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Splits ``range(len(weights))`` into ``world_size`` contiguous [begin, end) ranges whose weight sums are as
    even as a contiguous split allows (greedy on the prefix sum). Every frame lands in exactly one range."""
    n = len(weights)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if n == 0:
        return [(0, 0)] * world_size
    csum = np.cumsum(np.asarray(weights, dtype=np.int64))
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        cut = int(np.searchsorted(csum, target, side="left")) + 1
        cut = min(max(cut, bounds[-1]), n)
        # pick the closer of cut-1 / cut
        if cut - 1 > bounds[-1] and abs(int(csum[cut - 2]) - target) <= abs(int(csum[cut - 1]) - target):
            cut -= 1
        bounds.append(cut)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def my_shard(weights: Sequence[int], rank: int, world_size: int) -> Tuple[int, int]:
    return shard_bounds(weights, world_size)[rank]


def shard_bounds_native(weights: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """The C++ scheduler's split (ZSTDB200_shardBounds, what ZSTDB200_*BatchMulti use): must equal shard_bounds()."""
    import ctypes
    from . import _native
    n = len(weights)
    w = (ctypes.c_size_t * max(n, 1))(*[int(x) for x in weights])
    b = (ctypes.c_size_t * (world_size + 1))()
    _native.lib.ZSTDB200_shardBounds(n, w, world_size, b)
    return [(int(b[i]), int(b[i + 1])) for i in range(world_size)]
