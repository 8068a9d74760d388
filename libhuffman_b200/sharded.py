"""One job over several GPUs, one process per GPU: the driver of libhuffman_b200/shard.py.

`ShardedCodec` wraps a DeviceCodec (the C-ABI of include/huffman/b200.h) and a
torch.distributed group.  Tensors are plumbing here: device memory and the rendezvous.  It
works the same with CUDA tensors on the real library (nccl) and with CPU tensors on the
kernel-logic emulation (gloo; tests only).
"""
from __future__ import annotations

from . import shard
from .capi import DeviceCodec


class ShardedCodec:
    def __init__(self, lib, rank: int, world: int, device, device_index: int = -1, accept_1025: bool | None = None):
        import torch
        self.torch = torch
        self.lib, self.rank, self.world, self.device = lib, rank, world, device
        self.enc = DeviceCodec(lib, device_index)
        self.dec = DeviceCodec(lib, device_index, accept_1025=accept_1025)

    def close(self):
        self.enc.close()
        self.dec.close()

    # ---- small host-side exchanges (sizes, tuples) ------------------------------------------------

    def all_gather_ints(self, values: list[int]) -> list[list[int]]:
        torch = self.torch
        if self.world == 1:
            return [list(values)]
        import torch.distributed as dist
        t = torch.tensor(values, dtype=torch.int64, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]

    # ---- encode: my block range -> my slab ----------------------------------------------------------

    def encode_async(self, x, blocksize: int, out, stream: int = 0) -> None:
        self.enc.encode_async(x.data_ptr(), x.numel(), blocksize, out.data_ptr(), out.numel(), stream)

    def encode_finish(self) -> int:
        return self.enc.encode_finish()

    # ---- the one stream laid out by byte range ------------------------------------------------------

    def redistribute(self, slab, sizes: list[int], overlap: int):
        """Rank r ends up with stream bytes [cut_r rounded down to 16, cut_{r+1} + overlap).
        Returns (buffer, stream offset of buffer[0], cuts).  Bytes that already live on the rank
        stay there; only what lies around the cuts travels."""
        torch = self.torch
        offs = shard.slab_offsets(sizes)
        total = offs[-1]
        cuts = shard.stream_cuts(total, self.world)
        want = [((cuts[d] & ~15), min(total, cuts[d + 1] + overlap)) for d in range(self.world)]
        lo, hi = want[self.rank]
        if self.world == 1:
            return slab[: sizes[0]], 0, cuts
        import torch.distributed as dist
        # what I hold of everybody's range; my own share is copied locally, not through the exchange
        in_splits, send = [], []
        for d in range(self.world):
            a, b = max(want[d][0], offs[self.rank]), min(want[d][1], offs[self.rank + 1])
            n = max(0, b - a) if d != self.rank else 0
            in_splits.append(n)
            if n:
                send.append(slab[a - offs[self.rank]: b - offs[self.rank]])
        inp = torch.cat(send) if send else torch.empty(0, dtype=torch.uint8, device=self.device)
        pieces = [max(0, min(hi, offs[s + 1]) - max(lo, offs[s])) for s in range(self.world)]
        out_splits = [p if s != self.rank else 0 for s, p in enumerate(pieces)]
        got = torch.empty(sum(out_splits), dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(got, inp, out_splits, in_splits)
        buf = torch.empty(sum(pieces) + 64, dtype=torch.uint8, device=self.device)
        at = src = 0
        for s, p in enumerate(pieces):
            if s == self.rank:
                a = max(lo, offs[s]) - offs[s]
                buf[at:at + p] = slab[a:a + p]
            else:
                buf[at:at + p] = got[src:src + p]
                src += p
            at += p
        return buf[: sum(pieces)], lo, cuts

    # ---- decode: the blocks that start in my byte range ----------------------------------------------

    def decode_plan(self, buf, base: int, cuts: list[int], stream: int = 0) -> int:
        lo, hi = cuts[self.rank], cuts[self.rank + 1]
        return self.dec.decode_range_plan(buf.data_ptr(), buf.numel(), lo - base, hi - base, stream,
                                          start_is_block=self.rank == 0)

    def decode_async(self, buf, base: int, cuts: list[int], out, stream: int = 0) -> None:
        lo, hi = cuts[self.rank], cuts[self.rank + 1]
        self.dec.decode_range_async(buf.data_ptr(), buf.numel(), lo - base, hi - base, out.data_ptr(), out.numel(),
                                    stream, start_is_block=self.rank == 0)

    def decode_finish(self, base: int) -> tuple[int, int | None, int, int]:
        """(huf_error_t, first block offset or None, chain end, decoded bytes), stream offsets."""
        rc, first, end, n = self.dec.decode_range_finish()
        return rc, (None if first is None else first + base), end + base, n

    def validate(self, mine: tuple[int, int | None, int, int], cuts: list[int]):
        """Chain check across the ranks.  Returns (ok, chain end, output offsets per rank)."""
        rc, first, end, n = mine
        rows = self.all_gather_ints([rc, -1 if first is None else first, end, n])
        if any(r[0] != 0 for r in rows):
            return False, 0, []
        parts = [(None if r[1] < 0 else r[1], r[2], r[3]) for r in rows]
        return shard.validate_chain(parts, cuts)
