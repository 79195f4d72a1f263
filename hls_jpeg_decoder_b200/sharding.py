"""Multi-GPU sharding of a batch of independent images.

Images are independent units and there is no exchange step, so a batch is cut into one contiguous
image range per GPU, balanced by compressed bytes (the entropy kernel's time follows scan size).
No collective, no NVLink traffic: every rank decodes its own range with its own BatchDecoder.
"""
from __future__ import annotations


def shard_range(sizes: list[int], rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) image range of `rank`: contiguous, disjoint, covering, balanced by sum(sizes)."""
    n = len(sizes)
    if world <= 1:
        return 0, n
    total = sum(sizes)
    bounds = [0]
    acc, k = 0, 1
    for i, s in enumerate(sizes):
        acc += s
        while k < world and acc * world >= total * k:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world:
        bounds.append(n)
    bounds.append(n)
    bounds = [min(b, n) for b in bounds]
    for j in range(1, len(bounds)):
        bounds[j] = max(bounds[j], bounds[j - 1])
    return bounds[rank], bounds[rank + 1]
