"""Multi-GPU sharding of the block codec, one process per GPU.

Encode: blocks are self-contained (reference src/encoder.c:345,360-373 resets all state per
block), so rank r of N encodes the contiguous block range [r*B/N, (r+1)*B/N) into its own slab;
the ONE stream is the slabs in rank order, slab r starting at the exclusive scan of the slab
sizes.  Only the sizes are exchanged.

Decode: the stream has no block index (src/decoder.c:218-276 finds block k+1 by decoding block
k), so rank r takes the BYTE range [r*C/N, (r+1)*C/N) of the one stream plus an overlap for its
last block, finds the headers in it, decodes the blocks that start in it
(huf_b200_decode_range_async) and reports (first block offset, chain end, decoded bytes).  The
chain is validated across ranks on the host: range r must begin where the chain of the ranges
before it ended.  Decoded slabs concatenate in rank order.

No data-path collective: `torch.distributed` carries sizes, tuples and -- once, to lay the one
stream out by byte range the way a reader of a file would find it -- the few MiB around every
cut that live on the neighbouring rank.
"""
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


def stream_cuts(total: int, world: int) -> list[int]:
    """Byte ranges of the one stream for decode: rank r scans [cuts[r], cuts[r+1])."""
    return [total // world * r for r in range(world)] + [total]


def range_pieces(offs: list[int], lo: int, hi: int) -> list[tuple[int, int, int]]:
    """Which slab bytes make up stream bytes [lo, hi): (rank, from, to) in slab-local offsets."""
    out = []
    for r in range(len(offs) - 1):
        a, b = max(lo, offs[r]), min(hi, offs[r + 1])
        if a < b:
            out.append((r, a - offs[r], b - offs[r]))
    return out


def validate_chain(parts: list[tuple[int | None, int, int]], cuts: list[int]) -> tuple[bool, int, list[int]]:
    """Host-side chain check over per-rank (first block offset or None, chain end, decoded bytes),
    all in stream offsets.  Returns (ok, end of the chain, exclusive scan of the decoded sizes).
    A range that lies entirely inside a block of an earlier range contributes nothing."""
    expect, outs = 0, [0]
    for r, (first, end, n) in enumerate(parts):
        if r > 0 and expect >= cuts[r + 1]:
            outs.append(outs[-1])
            continue
        if first != expect:
            return False, expect, outs
        expect = end
        outs.append(outs[-1] + n)
    return expect == cuts[-1], expect, outs
