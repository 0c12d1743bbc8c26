"""Tuple-index sharding for the multi-GPU path (one process per GPU, no collective on the data path).

Tuples are independent, so rank r of w takes one contiguous block; block starts are multiples of 32 so
verdict-bitmap words never straddle ranks (the same rule the C library uses across a context's devices).
"""
from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    per = (((n + world - 1) // world) + 31) & ~31
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def bitmap_words(lo: int, hi: int) -> Tuple[int, int]:
    """word range [a, b) of the global verdict bitmap owned by the shard [lo, hi)"""
    return lo // 32, (hi + 31) // 32
