"""Multi-GPU sharding of the block codec: contiguous block ranges, no data-path collective.

Blocks are self-contained (reference src/encoder.c:345,360-373 resets all state per block), so
GPU g of G encodes blocks [g*B/G, (g+1)*B/G) into its own slab and the host concatenates the
slabs in rank order; slab g starts at the exclusive scan of the slab sizes.  Only sizes are
exchanged (a few bytes per rank, host side)."""
from __future__ import annotations


def block_range(nblocks: int, rank: int, world: int) -> tuple[int, int]:
    """Half-open block range of `rank`; ranges are contiguous, ordered and differ by <= 1."""
    base, extra = divmod(nblocks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def byte_range(length: int, blocksize: int, rank: int, world: int) -> tuple[int, int]:
    """Byte range of the uncompressed input that `rank` encodes."""
    if blocksize == 0:
        blocksize = length
    nblocks = (length + blocksize - 1) // blocksize if length else 0
    lo, hi = block_range(nblocks, rank, world)
    return min(lo * blocksize, length), min(hi * blocksize, length)


def slab_offsets(slab_sizes: list[int]) -> list[int]:
    """Exclusive scan of per-rank compressed sizes: where each slab starts in the stream."""
    out, run = [], 0
    for s in slab_sizes:
        out.append(run)
        run += s
    out.append(run)
    return out
