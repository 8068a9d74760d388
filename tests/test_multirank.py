"""N > 1 path on CPU: world_size-2 gloo.  Each rank encodes its contiguous block range (kernel
logic through the emulation library, test infrastructure), sizes are exchanged, slabs are
concatenated by the exclusive scan of the sizes; the result must equal the oracle's encoding
of the whole input, and each rank must decode its own slab back."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, tmp: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import torch
    import torch.distributed as dist

    import build_emu
    from libhuffman_b200 import datagen, shard
    from libhuffman_b200.capi import B200Lib
    from oracle import harness

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = B200Lib(build_emu.build())

    bs = 4096
    data = datagen.zipf(11 * bs + 123, 200, seed=3)      # 12 blocks, last one short
    lo, hi = shard.byte_range(len(data), bs, rank, world)
    rc, slab = emu.encode(data[lo:hi], bs)
    assert rc == 0

    sizes = [None] * world
    dist.all_gather_object(sizes, len(slab))
    offs = shard.slab_offsets(sizes)
    slabs = [None] * world
    dist.all_gather_object(slabs, slab)
    whole = b"".join(slabs)
    assert len(whole) == offs[-1]
    assert whole[offs[rank]:offs[rank + 1]] == slab
    assert whole == harness.oracle_encode(data, bs)

    # decode shards too: a slab is a valid stream of its own
    rc, back = emu.decode(slab)
    assert rc == 0 and back == data[lo:hi]
    parts = [None] * world
    dist.all_gather_object(parts, back)
    assert b"".join(parts) == data
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


def test_two_rank_block_range_sharding(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    from oracle import harness
    build_emu.build()
    harness.build()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_block_ranges_partition():
    from libhuffman_b200 import shard
    for nblocks in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 4, 8):
            r = [shard.block_range(nblocks, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nblocks
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard.byte_range(1000, 300, 1, 2) == (600, 1000)
    assert shard.slab_offsets([5, 7, 1]) == [0, 5, 12, 13]
