"""Sample-mode sharding helpers: every rank holds the same
contiguous block of rows of every coupled tensor and of Y."""

from __future__ import annotations


def row_block(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of rank ``rank``'s rows: blocks differ by at most one row,
    the first ``n_total % world`` ranks take the extra one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(arrays, rank: int, world: int):
    """Slice the sample mode of X (or each X of a list) and of Y alike."""
    if isinstance(arrays, (list, tuple)):
        return [shard_rows(a, rank, world) for a in arrays]
    lo, hi = row_block(arrays.shape[0], rank, world)
    return arrays[lo:hi]
