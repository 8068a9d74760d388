"""Kernel LOGIC on the CPU box: the unmodified CUDA sources run under tests/emu/cuda_emu.h (a
fiber SIMT emulator, test infrastructure) and are compared with the oracle bit for bit.  This
lane proves tie-breaks, bit offsets, byte ownership, header scan, speculative decode and error
codes before GPU time is spent; the `-m gpu` lane repeats the checks on the real library."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from cases import foreign_streams, small_cases
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec

CASES = small_cases()


@pytest.mark.parametrize("name,data,bs", CASES, ids=[c[0] for c in CASES])
def test_emu_encode_bit_exact(emu, harness, name, data, bs):
    rc, got = emu.encode(data, bs)
    assert rc == 0
    assert got == harness.oracle_encode(data, bs)


@pytest.mark.parametrize("name,data,bs", CASES, ids=[c[0] for c in CASES])
def test_emu_decode_matches_oracle(emu, harness, name, data, bs):
    stream = harness.oracle_encode(data, bs)
    rc_o, out_o, _ = harness.oracle_decode(stream)
    rc, got = emu.decode(stream)
    assert rc == rc_o          # strict mode: 256-symbol blocks are BTREE_OVERFLOW like the reference
    if rc == 0:
        assert got == data


def test_emu_golden_vectors(emu, golden):
    for v in golden["encode"]:
        rc, got = emu.encode(bytes.fromhex(v["input"]), v["blocksize"])
        assert rc == 0 and got == bytes.fromhex(v["stream"]), v["name"]
    for v in golden["decode"]:
        rc, got = emu.decode(bytes.fromhex(v["stream"]), v["length"])
        assert rc == v["rc"], v["name"]
        if rc == 0:
            assert got == bytes.fromhex(v["output"]), v["name"]


def test_emu_foreign_streams(emu, harness):
    tail = harness.oracle_encode(b"normal block after foreign ones", 0)
    for name, s in foreign_streams():
        for stream in (s, s + tail):
            rc_o, out_o, _ = harness.oracle_decode(stream)
            rc, got = emu.decode(stream)
            assert (rc, got) == (rc_o, out_o), name


def test_emu_lenient_1025_roundtrip(emu, harness):
    data = bytes(range(256)) * 5 + datagen.uniform(3000, 256, seed=3)
    stream = harness.oracle_encode(data, 1500)
    codec = DeviceCodec(emu, accept_1025=True)
    try:
        src = C.create_string_buffer(stream, len(stream) + 16)
        out = C.create_string_buffer(len(data) + 64)
        codec.decode_async(C.addressof(src), len(stream), len(stream), C.addressof(out), len(data) + 64)
        rc, n, used = codec.decode_finish()
        assert (rc, n, used) == (0, len(data), len(stream))
        assert out.raw[:n] == data
    finally:
        codec.close()


def test_emu_device_api_offsets_and_capacity(emu, harness):
    data = datagen.zipf(30000, 200, seed=12)
    bs = 4096
    want = harness.oracle_encode(data, bs)
    codec = DeviceCodec(emu)
    try:
        src = C.create_string_buffer(data, len(data))
        cap = codec.encode_bound(len(data), bs)
        dst = C.create_string_buffer(cap)
        codec.encode_async(C.addressof(src), len(data), bs, C.addressof(dst), cap)
        n = codec.encode_finish()
        assert dst.raw[:n] == want
        ptr, nb = codec.block_offsets()
        offs = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), (nb + 1,)).copy()
        assert nb == 8 and offs[0] == 0 and offs[-1] == n
        # every offset is a block header: orig_len field equals the block size
        for i in range(nb):
            o = int(offs[i])
            assert int.from_bytes(want[o:o + 8], "little") == min(bs, len(data) - i * bs)
        # too small an output buffer is reported, nothing is written out of bounds
        small = C.create_string_buffer(n // 2)
        codec.encode_async(C.addressof(src), len(data), bs, C.addressof(small), n // 2)
        with pytest.raises(Exception):
            codec.encode_finish()
    finally:
        codec.close()


def test_emu_corrupted_streams_error_parity(emu, harness):
    rng = np.random.default_rng(17)
    base = harness.oracle_encode(datagen.english_text(2500, seed=3), 700)
    for it in range(40):
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
        rc, got = emu.decode(s)
        assert rc == rc_o, (it, kind, rc, rc_o)
        if rc == 0:
            assert got == out_o


def test_emu_short_reader_and_length_semantics(emu, harness):
    data = datagen.english_text(1000, seed=2)
    # encoder: reader runs dry in the third block -> READ_WRITE after two whole blocks (Q11)
    rc, got = emu.encode(data[:700], 300, length=1000)
    assert rc == 3 and got == harness.oracle_encode(data[:600], 300)
    # decoder: `length` is only checked between blocks (src/decoder.c:218)
    stream = harness.oracle_encode(data, 400)
    rc, got = emu.decode(stream, length=5)
    assert (rc, got) == (0, data[:400])
    rc, got = emu.decode(stream + b"\x07\x07\x07")
    assert rc == 3 and got == data     # junk after the last whole block: READ_WRITE after the output


def _lanes(lib, stream, out_cap, accept_1025=False):
    """Decode through the device entry points; returns (rc, bytes, slow-lane block count)."""
    codec = DeviceCodec(lib, accept_1025=accept_1025)
    try:
        src = C.create_string_buffer(stream, len(stream) + 16)
        out = C.create_string_buffer(out_cap + 64)
        codec.decode_async(C.addressof(src), len(stream), len(stream), C.addressof(out), out_cap + 64)
        rc, n, used = codec.decode_finish()
        return rc, out.raw[:n], codec.slow_blocks()
    finally:
        codec.close()


def _deep_blocks(harness, data, bs, reach=32):
    """Blocks of oracle_encode(data, bs) whose longest code word exceeds the table reach."""
    deep = 0
    step = bs if bs else len(data)
    for i in range(0, len(data), step):
        blk = data[i:i + step]
        lens, _, _ = harness.oracle_codebook([blk.count(bytes([b])) for b in range(256)])
        deep += max(lens) > reach
    return deep


def test_emu_fast_lane_takes_encoder_shaped_blocks(emu, harness):
    """Blocks with the reference encoder's tree shape must be decoded by the fast lane (codes
    beyond the 13-bit table go through its long-code records); only code words beyond 32 bits
    (and, in strict mode, 1025-element trees) may take the general lane."""
    for name, data, bs in CASES:
        stream = harness.oracle_encode(data, bs)
        rc, got, slow = _lanes(emu, stream, len(data), accept_1025=True)
        assert (rc, got) == (0, data), name
        assert slow == _deep_blocks(harness, data, bs), name
    # strict mode: the 1025-element tree is refused by the general lane like the reference does
    data = bytes(range(256)) * 3
    rc, got, slow = _lanes(emu, harness.oracle_encode(data, 0), len(data))
    assert rc == 5 and slow == 1


def test_emu_corrupted_large_block_error_parity(emu, harness):
    """Same as the GPU lane's large-block corruption test, a few cases: damaged 64 KiB blocks must
    leave the fast lane and get the oracle's error code / bytes from the general lane."""
    rng = np.random.default_rng(23)
    data = datagen.zipf(2 * 65536 + 77, 255, seed=5)
    base = harness.oracle_encode(data, 65536)
    for it in range(6):
        s = bytearray(base)
        if it % 3 == 0:
            s = s[: int(rng.integers(len(s) // 2, len(s)))]
        elif it % 3 == 1:
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        else:
            pos = int(rng.integers(0, len(s) - 64))
            s[pos:pos + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        s = bytes(s)
        rc_o, out_o, _ = harness.oracle_decode(s)
        rc, got = emu.decode(s)
        assert rc == rc_o, (it, rc, rc_o)
        if rc == 0:
            assert got == out_o, it


def test_emu_decode_with_block_index_hint(emu, harness):
    """the encoder's block-offset array as a decode-side index.  With it the
    header scan is skipped; a wrong index must not change the result (chain check + rescan)."""
    data = datagen.zipf(50000, 200, seed=12)
    bs = 4096
    want = harness.oracle_encode(data, bs)
    enc = DeviceCodec(emu)
    dec = DeviceCodec(emu)
    try:
        src = C.create_string_buffer(data, len(data))
        cap = enc.encode_bound(len(data), bs)
        comp = C.create_string_buffer(cap + 16)
        enc.encode_async(C.addressof(src), len(data), bs, C.addressof(comp), cap)
        n = enc.encode_finish()
        assert comp.raw[:n] == want
        ptr, nb = enc.block_offsets()
        offs = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), (nb + 1,)).copy()
        out = C.create_string_buffer(len(data) + 64)
        # the encoder's own index: no k_find launch
        dec.decode_hint_offsets(ptr, nb)
        dec.decode_async(C.addressof(comp), n, n, C.addressof(out), len(data) + 64)
        hinted_launches = dec.launches()
        assert dec.decode_finish() == (0, len(data), n) and out.raw[:len(data)] == data
        # without the hint one more kernel family runs (the scan)
        dec.decode_async(C.addressof(comp), n, n, C.addressof(out), len(data) + 64)
        assert dec.launches() > hinted_launches
        assert dec.decode_finish() == (0, len(data), n)
        # a damaged index (one offset off by 3, one block missing) costs a rescan, not correctness
        bad = offs[:nb].copy()
        bad[3] += 3
        bad = np.delete(bad, 7)
        badbuf = (C.c_uint64 * len(bad))(*bad.tolist())
        out2 = C.create_string_buffer(len(data) + 64)
        dec.decode_hint_offsets(C.addressof(badbuf), len(bad))
        dec.decode_async(C.addressof(comp), n, n, C.addressof(out2), len(data) + 64)
        assert dec.decode_finish() == (0, len(data), n) and out2.raw[:len(data)] == data
    finally:
        enc.close()
        dec.close()
