"""Parity tests proper: the real CUDA library on a B200, called through its C-ABI, against the
oracle on the same seeded inputs (bit-exact: this is byte/integer work), against the committed
golden vectors, and at BASELINE.json's full size through size-independent properties."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from cases import foreign_streams, small_cases
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec

pytestmark = pytest.mark.gpu

CASES = small_cases()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the gpu lane needs a CUDA device"
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def codec(lib, torch_cuda):
    c = DeviceCodec(lib, 0)
    yield c
    c.close()


def dev_encode(torch, codec, data_t, blocksize):
    """Device-resident encode of a uint8 CUDA tensor; returns (stream tensor, block offsets)."""
    n = data_t.numel()
    cap = codec.encode_bound(n, blocksize)
    out = torch.empty(cap, dtype=torch.uint8, device=data_t.device)
    stream = torch.cuda.current_stream().cuda_stream
    codec.encode_async(data_t.data_ptr(), n, blocksize, out.data_ptr(), cap, stream)
    size = codec.encode_finish()
    ptr, nb = codec.block_offsets()
    host = np.empty(nb + 1, dtype=np.uint64)   # the offsets array lives in device memory
    codec.lib.check(codec.lib.dll.huf_b200_copy_d2h(host.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), 8 * (nb + 1)),
                    "copy offsets")
    return out[:size], host


def dev_decode(torch, codec, stream_t, out_len, length=None):
    out = torch.empty(out_len + 64, dtype=torch.uint8, device=stream_t.device)
    st = torch.cuda.current_stream().cuda_stream
    n = stream_t.numel()
    codec.decode_async(stream_t.data_ptr(), n, n if length is None else length, out.data_ptr(), out_len + 64, st)
    rc, produced, used = codec.decode_finish()
    return rc, out[:produced], used


@pytest.mark.parametrize("name,data,bs", CASES, ids=[c[0] for c in CASES])
def test_c_api_encode_bit_exact(lib, harness, name, data, bs):
    """huf_encode over memory streams, exactly as reference test/encode_test.c drives it."""
    rc, got = lib.encode(data, bs)
    assert rc == 0
    assert got == harness.oracle_encode(data, bs)


@pytest.mark.parametrize("name,data,bs", CASES, ids=[c[0] for c in CASES])
def test_c_api_decode_matches_oracle(lib, harness, name, data, bs):
    stream = harness.oracle_encode(data, bs)
    rc_o, out_o, _ = harness.oracle_decode(stream)
    rc, got = lib.decode(stream)
    assert rc == rc_o
    if rc == 0:
        assert got == data


def test_golden_vectors(lib, golden):
    for v in golden["encode"]:
        rc, got = lib.encode(bytes.fromhex(v["input"]), v["blocksize"])
        assert rc == 0 and got == bytes.fromhex(v["stream"]), v["name"]
    for v in golden["decode"]:
        rc, got = lib.decode(bytes.fromhex(v["stream"]), v["length"])
        assert rc == v["rc"], v["name"]
        if rc == 0:
            assert got == bytes.fromhex(v["output"]), v["name"]


def test_reference_encode_test_roundtrip_with_buffers(lib):
    """reference test/encode_test.c:48-94: blocksize 0, 128-byte bufio hints, decode writes back
    into the input stream and is read from there."""
    from libhuffman_b200.capi import Config
    with lib.memstream(128) as inp, lib.memstream(2048) as out:
        inp.write(b"0123456789")
        cfg = Config(length=10, reader_buffer_size=128, writer_buffer_size=128, reader=inp.rw, writer=out.rw)
        assert lib.dll.huf_encode(C.byref(cfg)) == 0
        n = len(out)
        assert n == 98
        cfg.reader, cfg.writer, cfg.length = out.rw, inp.rw, n
        assert lib.dll.huf_decode(C.byref(cfg)) == 0
        assert inp.read(10) == b"0123456789"


def test_foreign_streams(lib, harness):
    tail = harness.oracle_encode(b"normal block after foreign ones", 0)
    for name, s in foreign_streams():
        for stream in (s, s + tail):
            rc_o, out_o, _ = harness.oracle_decode(stream)
            rc, got = lib.decode(stream)
            assert (rc, got) == (rc_o, out_o), name


def test_corrupted_streams_error_parity(lib, harness):
    rng = np.random.default_rng(17)
    base = harness.oracle_encode(datagen.english_text(20000, seed=3), 3000)
    for it in range(60):
        s = bytearray(base)
        kind = it % 3
        if kind == 0:
            s = s[: int(rng.integers(1, len(s)))]
        elif kind == 1:
            for _ in range(2):
                s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        else:
            pos = int(rng.integers(0, len(s) - 4))
            s[pos:pos + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        s = bytes(s)
        rc_o, out_o, _ = harness.oracle_decode(s)
        rc, got = lib.decode(s)
        assert rc == rc_o, (it, kind, rc, rc_o)
        if rc == 0:
            assert got == out_o


@pytest.mark.parametrize("shape", ["zipf255", "geometric"])
def test_corrupted_large_blocks_error_parity(lib, harness, shape):
    """Bit flips, overwrites and truncation in streams of 64 KiB blocks (several chunks per block
    in the fast lane, re-speculation for the geometric shape): the error code, and for successful
    decodes the bytes, must be the oracle's; the fast lane has to hand every damaged block to the
    general lane.  A flip inside a payload usually still decodes (to different bytes)."""
    rng = np.random.default_rng(23)
    n = 5 * 65536 + 1234
    data = datagen.zipf(n, 255, seed=5) if shape == "zipf255" else datagen.geometric(n, seed=5)
    base = harness.oracle_encode(data, 65536)
    for it in range(40):
        s = bytearray(base)
        kind = it % 4
        if kind == 0:
            s = s[: int(rng.integers(len(s) // 2, len(s)))]
        elif kind == 1:
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            pos = int(rng.integers(0, len(s) - 64))
            s[pos:pos + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        else:
            pos = int(rng.integers(0, len(s) - 4096))
            s[pos:pos + 4096] = bytes(4096)
        s = bytes(s)
        rc_o, out_o, _ = harness.oracle_decode(s)
        rc, got = lib.decode(s)
        assert rc == rc_o, (it, kind, rc, rc_o)
        if rc == 0:
            assert got == out_o, (it, kind)


def test_short_reader_and_length_semantics(lib, harness):
    data = datagen.english_text(1000, seed=2)
    rc, got = lib.encode(data[:700], 300, length=1000)
    assert rc == 3 and got == harness.oracle_encode(data[:600], 300)
    stream = harness.oracle_encode(data, 400)
    rc, got = lib.decode(stream, length=5)
    assert (rc, got) == (0, data[:400])
    rc, got = lib.decode(stream + b"\x07\x07\x07")
    assert rc == 3 and got == data


@pytest.mark.parametrize("shape", ["english", "zipf256", "zipf255", "uniform", "fibonacci", "geometric"])
def test_medium_inputs_vs_oracle(lib, harness, shape):
    """Config 1 (1 MiB English-like, 64 KiB blocks) and 8 MiB of every other named shape."""
    n = 1 << 20 if shape == "english" else 8 << 20
    data = {
        "english": lambda: datagen.english_text(n, seed=1),
        "zipf256": lambda: datagen.zipf(n, 256, seed=2),
        "zipf255": lambda: datagen.zipf(n, 255, seed=2),
        "uniform": lambda: datagen.uniform(n, 256, seed=3),
        "fibonacci": lambda: datagen.fibonacci(n, 65536, seed=4),
        "geometric": lambda: datagen.geometric(n, seed=4),
    }[shape]()
    want = harness.oracle_encode(data, 65536)
    rc, got = lib.encode(data, 65536)
    assert rc == 0 and got == want
    rc_o, _, _ = harness.oracle_decode(want[: 1 << 20])  # strictness of the first blocks
    rc, back = lib.decode(want)
    if shape in ("zipf256", "uniform"):
        assert rc == 5            # Q2: 1025-element trees are rejected exactly like the reference
    else:
        assert rc == 0 and back == data


@pytest.mark.parametrize("bs", [4096, 16384, 262144, 1 << 20, 0])
def test_blocksize_sweep_vs_oracle(lib, harness, bs):
    data = datagen.zipf(3 << 20, 255, seed=5)
    want = harness.oracle_encode(data, bs)
    rc, got = lib.encode(data, bs)
    assert rc == 0 and got == want
    rc, back = lib.decode(want)
    assert rc == 0 and back == data


def test_deep_tree_1mib_block(lib, harness):
    """Config 3(iii): Fibonacci counts in a 1 MiB block -> 28-bit code words (64-bit table path)."""
    data = datagen.fibonacci(2 << 20, 1 << 20, seed=4)
    want = harness.oracle_encode(data, 1 << 20)
    rc, got = lib.encode(data, 1 << 20)
    assert rc == 0 and got == want
    rc, back = lib.decode(want)
    assert rc == 0 and back == data


def test_reference_decodes_gpu_output_and_vice_versa(lib, harness):
    if not harness.reference_available():
        pytest.skip("oracle/_ref not present")
    ref = harness.reference()
    data = datagen.zipf(2 << 20, 255, seed=6)
    rc, gpu_stream = lib.encode(data, 65536)
    assert rc == 0
    rc, ref_stream = ref.encode(data, 65536)
    assert rc == 0 and gpu_stream == ref_stream
    rc, back = ref.decode(gpu_stream)
    assert rc == 0 and back == data
    rc, back = lib.decode(ref_stream)
    assert rc == 0 and back == data


def test_device_api_full_size_properties(lib, harness, torch_cuda, codec):
    """1 GiB Zipf(1.1), 64 KiB blocks (BASELINE configs[1]): sampled blocks equal the oracle's
    encoding of the same block, offsets are consistent, and decode(encode(x)) == x on device."""
    torch = torch_cuda
    n = 1 << 30
    bs = 65536
    x = datagen.zipf_torch(n, "cuda", 256, seed=2)
    stream, offs = dev_encode(torch, codec, x, bs)
    nb = n // bs
    assert len(offs) == nb + 1 and offs[0] == 0 and int(offs[-1]) == stream.numel()
    assert np.all(np.diff(offs.astype(np.int64)) > 0)
    rng = np.random.default_rng(1)
    for b in [0, 1, nb - 1, *rng.integers(0, nb, 24).tolist()]:
        block = x[b * bs:(b + 1) * bs].cpu().numpy().tobytes()
        got = stream[int(offs[b]):int(offs[b + 1])].cpu().numpy().tobytes()
        assert got == harness.oracle_encode(block, 0), f"block {b}"
    # strict decode refuses the 1025-element trees like the reference (Q2) ...
    rc, _, _ = dev_decode(torch, codec, stream, n)
    assert rc == 5
    # ... the opt-in mode round-trips
    codec.set_accept_1025(True)
    try:
        rc, back, used = dev_decode(torch, codec, stream, n)
        assert rc == 0 and used == stream.numel() and back.numel() == n
        assert torch.equal(back, x)
    finally:
        codec.set_accept_1025(False)


def test_device_api_255_symbol_roundtrip_strict(lib, torch_cuda, codec):
    torch = torch_cuda
    n = 256 << 20
    x = datagen.zipf_torch(n, "cuda", 255, seed=7)
    stream, offs = dev_encode(torch, codec, x, 65536)
    rc, back, used = dev_decode(torch, codec, stream, n)
    assert rc == 0 and used == stream.numel()
    assert torch.equal(back, x)


def test_encode_is_idempotent_and_block_local(lib, torch_cuda, codec):
    """Blocks are independent: encoding a contiguous block range alone gives the same bytes as
    the matching slice of the whole stream (the multi-GPU sharding property)."""
    torch = torch_cuda
    bs = 65536
    x = datagen.zipf_torch(64 << 20, "cuda", 256, seed=9)
    whole, offs = dev_encode(torch, codec, x, bs)
    whole = whole.clone()
    half = (32 << 20) // bs
    first, _ = dev_encode(torch, codec, x[: 32 << 20], bs)
    first = first.clone()
    second, _ = dev_encode(torch, codec, x[32 << 20:], bs)
    assert torch.equal(torch.cat([first, second]), whole)
    assert first.numel() == int(offs[half])


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [4096, 65536, 1 << 20])
@pytest.mark.parametrize("shape", ["english", "zipf255", "fibonacci", "geometric"])
def test_shape_blocksize_matrix(lib, harness, shape, bs):
    """Every named shape at small, headline and large blocks: long code words (English text at
    4 KiB blocks reaches 14 bits, Fibonacci counts 22-28 bits), many chunks per block, blocks
    smaller than one chunk.  8 MiB of Fibonacci data at 64 KiB blocks once exposed a divergent
    barrier in the fast decode lane, hence the size."""
    n = 8 << 20
    data = {
        "english": lambda: datagen.english_text(n, seed=1),
        "zipf255": lambda: datagen.zipf(n, 255, seed=2),
        "fibonacci": lambda: datagen.fibonacci(n, 65536, seed=4),
        "geometric": lambda: datagen.geometric(n, seed=4),
    }[shape]()
    want = harness.oracle_encode(data, bs)
    rc, got = lib.encode(data, bs)
    assert rc == 0 and got == want
    rc, back = lib.decode(want)
    assert rc == 0 and back == data


def test_decode_with_block_index_hint(lib, torch_cuda, codec):
    """SURVEY.md §8(f)4 on the device: the encoder's offset array as decode-side index (no header
    scan), and a damaged index that must not change the result."""
    torch = torch_cuda
    n, bs = 8 << 20, 65536
    x = datagen.zipf_torch(n, torch.device("cuda", 0), 255, seed=9)
    cap = codec.encode_bound(n, bs)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    codec.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, st)
    c = codec.encode_finish()
    ptr, nb = codec.block_offsets()
    assert nb == n // bs
    dec = DeviceCodec(lib, 0)
    try:
        back = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        dec.decode_hint_offsets(ptr, nb)
        dec.decode_async(comp.data_ptr(), c, c, back.data_ptr(), n + 64, st)
        hinted = dec.launches()
        assert dec.decode_finish() == (0, n, c) and torch.equal(back[:n], x)
        dec.decode_async(comp.data_ptr(), c, c, back.data_ptr(), n + 64, st)
        assert dec.launches() > hinted
        assert dec.decode_finish() == (0, n, c)
        # copy the device offsets, damage them, feed them back
        host = np.zeros(nb, dtype=np.uint64)
        lib.check(lib.dll.huf_b200_copy_d2h(host.ctypes.data, ptr, 8 * nb), "copy_d2h")
        host[5] += 7
        host = np.delete(host, 11)
        bad = torch.from_numpy(host.astype(np.int64)).cuda()
        back.zero_()
        dec.decode_hint_offsets(bad.data_ptr(), len(host))
        dec.decode_async(comp.data_ptr(), c, c, back.data_ptr(), n + 64, st)
        assert dec.decode_finish() == (0, n, c) and torch.equal(back[:n], x)
    finally:
        dec.close()
