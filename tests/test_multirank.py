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


def _sharded_worker(rank: int, world: int, port: int, tmp: str):
    """The one-process-per-GPU driver (libhuffman_b200/sharded.py) on CPU tensors: block ranges
    in, one stream out, the stream laid out by byte range, range decode, chain check."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import numpy as np
    import torch
    import torch.distributed as dist

    import build_emu
    from libhuffman_b200 import datagen, shard
    from libhuffman_b200.capi import B200Lib
    from libhuffman_b200.sharded import ShardedCodec
    from oracle import harness

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = B200Lib(build_emu.build())
    sc = ShardedCodec(emu, rank, world, torch.device("cpu"))
    for bs, n in ((2048, 23 * 2048 + 77), (4096, 6 * 4096)):
        data = datagen.zipf(n, 150, seed=9)
        whole = harness.oracle_encode(data, bs)
        lo, hi = shard.byte_range(n, bs, rank, world)
        x = torch.frombuffer(bytearray(data[lo:hi]), dtype=torch.uint8)
        comp = torch.empty(sc.enc.encode_bound(hi - lo, bs), dtype=torch.uint8)
        sc.encode_async(x, bs, comp)
        size = sc.encode_finish()
        sizes = [r[0] for r in sc.all_gather_ints([size])]
        offs = shard.slab_offsets(sizes)
        assert offs[-1] == len(whole)
        assert comp[:size].numpy().tobytes() == whole[offs[rank]:offs[rank + 1]]
        # the one stream by byte range: a small overlap on purpose (a block is ~2-4 KB here)
        buf, base, cuts = sc.redistribute(comp, sizes, overlap=6000)
        want_lo, want_hi = cuts[rank] & ~15, min(len(whole), cuts[rank + 1] + 6000)
        assert base == want_lo and buf.numpy().tobytes() == whole[want_lo:want_hi]
        est = sc.decode_plan(buf, base, cuts)
        out = torch.empty(est + 64, dtype=torch.uint8)
        sc.decode_async(buf, base, cuts, out)
        mine = sc.decode_finish(base)
        ok, end, outs = sc.validate(mine, cuts)
        assert ok and end == len(whole) and outs[-1] == n, (ok, end, outs)
        assert out[:mine[3]].numpy().tobytes() == data[outs[rank]:outs[rank + 1]]
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"sharded{rank}").write_text("ok")


def test_two_rank_sharded_codec(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    from oracle import harness
    build_emu.build()
    harness.build()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "sharded0").exists() and (tmp_path / "sharded1").exists()


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


_MULTI_DEVICE_SCRIPT = r'''
import sys
sys.path.insert(0, "{root}")
sys.path.insert(0, "{root}/tests")
sys.path.insert(0, "{root}/tests/emu")
import numpy as np
import build_emu
from cases import foreign_streams
from libhuffman_b200 import datagen
from libhuffman_b200.capi import B200Lib
from oracle import harness
harness.build()
lib = B200Lib(build_emu.build())
rng = np.random.default_rng(5)
for n, bs in ((11 * 4096 + 123, 4096), (9000, 1000), (30000, 0), (50000, 20000)):
    data = datagen.zipf(n, 200, seed=3)
    want = harness.oracle_encode(data, bs)
    rc, got = lib.encode(data, bs)
    assert rc == 0 and got == want, ("encode", n, bs)
    rc, back = lib.decode(want)
    assert (rc, back) == (0, data), ("decode", n, bs)
    # `length` in the middle of the stream: whole blocks up to the one that starts before it
    rc_o, out_o, _ = harness.oracle_decode(want, len(want) // 2)
    rc, back = lib.decode(want, length=len(want) // 2)
    assert (rc, back) == (rc_o, out_o), ("length", n, bs)
    # damaged streams: the error code and the output before the failing block are the oracle's
    for it in range(6):
        s = bytearray(want)
        if it % 3 == 0:
            s = s[: int(rng.integers(1, len(s)))]
        elif it % 3 == 1:
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        else:
            pos = int(rng.integers(0, max(1, len(s) - 8)))
            s[pos:pos + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
        s = bytes(s)
        rc_o, out_o, _ = harness.oracle_decode(s)
        rc, back = lib.decode(s)
        assert rc == rc_o and (rc != 0 or back == out_o), ("damaged", n, bs, it, rc, rc_o)
# headers the scan does not recognise: the seams do not validate, the serial lane takes over
tail = harness.oracle_encode(datagen.zipf(9000, 100, seed=4), 1500)
for name, s in foreign_streams():
    for stream in (s + tail, tail + s + tail):
        rc_o, out_o, _ = harness.oracle_decode(stream)
        rc, got = lib.decode(stream)
        assert (rc, got) == (rc_o, out_o), name
print("multi-device ok")
'''


def test_several_devices_in_one_call(tmp_path):
    """HUF_B200_DEVICES: huf_encode splits the block range, huf_decode the byte range of the one
    stream over the listed devices (three contexts on the emulated device here; the -m gpu lane
    repeats it on two real GPUs when the box has them).  Results must not depend on the split."""
    import subprocess
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    build_emu.build()
    env = dict(os.environ, HUF_B200_DEVICES="0,0,0", HUF_B200_MULTI_MIN="0")
    proc = subprocess.run([sys.executable, "-c", _MULTI_DEVICE_SCRIPT.format(root=ROOT)], env=env,
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and "multi-device ok" in proc.stdout, proc.stdout + proc.stderr
