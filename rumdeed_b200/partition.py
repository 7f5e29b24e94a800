"""i-range partition of the acceleration evaluation across the GPUs of one box.

Rows are split into equal contiguous chunks of ceil(n / world); the store capacity is padded to
chunk * world so that the (3, capacity) acceleration buffer can be all-gathered in place with
equal-sized slices.  Global indices are preserved (the image-charge roles depend on them).
"""
from __future__ import annotations


def row_partition(n: int, world: int, rank: int):
    """Returns (chunk, i_begin, i_end, padded_capacity) for `rank` of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    chunk = (n + world - 1) // world if n > 0 else 0
    i0 = min(n, rank * chunk)
    i1 = min(n, (rank + 1) * chunk)
    return chunk, i0, i1, max(1, chunk * world)
