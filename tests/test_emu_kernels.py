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


# ---- host lanes (huf_b200_encode_host / huf_b200_decode_host): spans, stage threads, streams ----

class _CallbackStream:
    """A user-supplied huf_read_writer_t (legal per reference include/huffman/io.h:11-21): read
    hands out at most `chunk` bytes per call, write collects."""

    def __init__(self, data: bytes = b"", chunk: int = 1000):
        from libhuffman_b200.capi import READ_FN, WRITE_FN, ReadWriter
        self.src = data
        self.pos = 0
        self.chunk = chunk
        self.out = bytearray()

        def _read(_stream, buf, count_p):
            n = min(count_p[0], self.chunk, len(self.src) - self.pos)
            C.memmove(buf, self.src[self.pos:self.pos + n], n)
            self.pos += n
            count_p[0] = n
            return 0

        def _write(_stream, buf, count):
            self.out += C.string_at(buf, count)
            return 0

        self._r, self._w = READ_FN(_read), WRITE_FN(_write)
        self.rw = ReadWriter(None, self._w, self._r)


def _codec_call(lib, fn, length, blocksize, reader, writer):
    from libhuffman_b200.capi import Config
    cfg = Config(length=length, blocksize=blocksize, reader=reader, writer=writer)
    return getattr(lib.dll, fn)(C.byref(cfg))


@pytest.fixture
def small_spans(monkeypatch):
    """Many spans out of a small input: every stage thread and slot hand-over of the host lanes runs."""
    monkeypatch.setenv("HUF_B200_SPAN_BYTES", "8192")


def test_emu_host_lane_many_spans(emu, harness, small_spans):
    data = datagen.zipf(70000 + 123, 200, seed=31)
    for bs in (4096, 1000, 30000):   # 2 blocks per span / 8 per span / a block larger than a span
        want = harness.oracle_encode(data, bs)
        rc, got = emu.encode(data, bs)
        assert rc == 0 and got == want, bs
        rc, back = emu.decode(want)
        assert (rc, back) == (0, data), bs
    # the reader runs dry inside the 6th span: whole blocks before it, then READ_WRITE (Q11)
    rc, got = emu.encode(data[:45000], 4096, length=70000)
    assert rc == 3 and got == harness.oracle_encode(data[:45000 // 4096 * 4096], 4096)
    # errors in a late span keep the output of the blocks before them (src/decoder.c:218-276)
    stream = harness.oracle_encode(data, 4096)
    rc_o, out_o, _ = harness.oracle_decode(stream[:-700])
    rc, got = emu.decode(stream[:-700])
    assert rc == rc_o == 3 and got == out_o
    bad = bytearray(stream)
    bad[len(bad) * 3 // 4] ^= 0x10
    rc_o, out_o, _ = harness.oracle_decode(bytes(bad))
    rc, got = emu.decode(bytes(bad))
    assert rc == rc_o and (rc != 0 or got == out_o)


def test_emu_host_lane_callback_and_fd_streams(emu, harness, small_spans, tmp_path):
    """Readers/writers that are not huf_memopen streams go through the pull/push side of the lanes."""
    data = datagen.english_text(50000, seed=8)
    bs = 3000
    want = harness.oracle_encode(data, bs)
    src, dst = _CallbackStream(data, chunk=777), _CallbackStream()
    assert _codec_call(emu, "huf_encode", len(data), bs, C.pointer(src.rw), C.pointer(dst.rw)) == 0
    assert bytes(dst.out) == want
    # decode: the reader holds more than `length`; the last block continues past it and the bytes
    # behind that block stay unread (the reference pulls what a block needs, src/decoder.c:218-261)
    src, dst = _CallbackStream(want + b"trailing", chunk=5000), _CallbackStream()
    assert _codec_call(emu, "huf_decode", len(want) - 100, 0, C.pointer(src.rw), C.pointer(dst.rw)) == 0
    assert bytes(dst.out) == data
    # fd streams (huf_fdopen), file to file
    from libhuffman_b200.capi import ReadWriter
    import os
    fin, fmid, fout = (str(tmp_path / n) for n in ("in", "mid", "out"))
    open(fin, "wb").write(data)
    emu.dll.huf_fdopen.argtypes = [C.POINTER(C.POINTER(ReadWriter)), C.c_int]
    for a, b, fn, length, blk in ((fin, fmid, "huf_encode", len(data), bs), (fmid, fout, "huf_decode", len(want), 0)):
        fa, fb = os.open(a, os.O_RDONLY), os.open(b, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
        ra, rb = C.POINTER(ReadWriter)(), C.POINTER(ReadWriter)()
        assert emu.dll.huf_fdopen(C.byref(ra), fa) == 0 and emu.dll.huf_fdopen(C.byref(rb), fb) == 0
        assert _codec_call(emu, fn, length, blk, ra, rb) == 0
        emu.dll.huf_fdclose(C.byref(ra))
        emu.dll.huf_fdclose(C.byref(rb))
        os.close(fa)
        os.close(fb)
    assert open(fmid, "rb").read() == want and open(fout, "rb").read() == data


def test_emu_decode_length_beyond_data(emu, harness):
    """`length` larger than the readable bytes with the data ending on a block
    boundary, and an empty reader, used to spin in the restart loop; the reference fails on the
    read of the next header with HUF_ERROR_READ_WRITE (src/decoder.c:220-229)."""
    data = datagen.english_text(1000, seed=2)
    stream = harness.oracle_encode(data, 400)
    rc_o, out_o, _ = harness.oracle_decode(stream, len(stream) + 10)
    rc, got = emu.decode(stream, length=len(stream) + 10)
    assert (rc, got) == (rc_o, out_o) == (3, data)
    rc, got = emu.decode(b"", length=10)
    assert (rc, got) == (3, b"")
    # the second decompress() of one HuffmanDecompressor: length counts consumed bytes again (P2)
    with emu.memstream(len(stream)) as src, emu.memstream(64) as dst:
        src.write(stream)
        assert _codec_call(emu, "huf_decode", len(stream), 0, src.rw, dst.rw) == 0
        assert _codec_call(emu, "huf_decode", len(stream), 0, src.rw, dst.rw) == 3
        assert dst.getvalue() == data


def test_emu_failed_async_leaves_context_usable(emu, harness):
    """an *_async call that fails before everything is enqueued must not
    leave the context pending."""
    data = datagen.zipf(9000, 100, seed=3)
    codec = DeviceCodec(emu)
    try:
        src = C.create_string_buffer(data, len(data))
        cap = codec.encode_bound(len(data), 1 << 40)
        dst = C.create_string_buffer(cap)
        # nspb does not fit 32 bits: INVALID_ARGUMENT from inside encode_async
        with pytest.raises(Exception):
            codec.encode_async(C.addressof(src), len(data), 1 << 50, C.addressof(dst), cap)
        out = C.create_string_buffer(len(data) + 64)
        stream = harness.oracle_encode(data, 4096)
        sbuf = C.create_string_buffer(stream, len(stream) + 16)
        codec.decode_async(C.addressof(sbuf), len(stream), len(stream), C.addressof(out), len(data) + 64)
        assert codec.decode_finish() == (0, len(data), len(stream))
        codec.encode_async(C.addressof(src), len(data), 4096, C.addressof(dst), cap)
        assert dst.raw[:codec.encode_finish()] == stream
    finally:
        codec.close()


def _range_decode_all(lib, stream, cuts, out_cap):
    """Decode `stream` as byte ranges cut at `cuts` (the multi-GPU decode of SURVEY.md §8(e) on one
    device): every range decodes the blocks that start in it; returns the stitched bytes after
    checking that each range begins where its predecessor's chain ended."""
    codec = DeviceCodec(lib)
    try:
        src = C.create_string_buffer(stream, len(stream) + 16)
        bounds = [0, *cuts, len(stream)]
        out, expect_first = b"", 0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            buf = C.create_string_buffer(out_cap + 64)
            codec.decode_range_async(C.addressof(src), len(stream), lo, hi, C.addressof(buf), out_cap + 64)
            rc, first, end, n = codec.decode_range_finish()
            assert rc == 0, (lo, hi, rc)
            if first is None:
                assert n == 0 and end == lo   # no block starts in this range (a block spans it)
                continue
            assert first == expect_first, (lo, hi, first, expect_first)
            expect_first = end
            out += buf.raw[:n]
        assert expect_first == len(stream)
        return out
    finally:
        codec.close()


def test_emu_decode_by_byte_ranges(emu, harness):
    data = datagen.zipf(12000, 180, seed=14)
    stream = harness.oracle_encode(data, 2048)          # 6 blocks of ~2 KB
    n = len(stream)
    offs = _block_offsets(stream)
    for cuts in ([n // 2], [1, n // 3, n - 1], [n // 4, n // 4 + 100, n // 4 + 200],
                 [offs[2], offs[4]]):                   # (the last: cuts exactly on block boundaries)
        assert _range_decode_all(emu, stream, cuts, len(data)) == data, cuts
    big = harness.oracle_encode(data[:5000], 0)         # one block: the ranges behind the first are empty
    assert _range_decode_all(emu, big, [len(big) // 3, len(big) // 2], 5000) == data[:5000]


def _block_offsets(stream):
    """Walk the block chain of a valid stream with the oracle (test helper)."""
    from oracle import harness as h
    offs, at = [], 0
    while at < len(stream):
        offs.append(at)
        rc, _, used = h.oracle_decode(stream[at:], 1)
        assert rc == 0
        at += used
    return offs


def test_emu_header_scan_interior_chunks(emu, harness):
    """k_find streams chunks that lie entirely inside the scanned range through a specialised loop
    (no per-row bounds tests, exact masks only for rows whose prefilter fired).  A stream of a few
    chunks with headers at many phases, plus payload bytes that look like headers."""
    rng = np.random.default_rng(3)
    data = datagen.zipf(150000, 200, seed=4)
    for bs in (3001, 20000):
        stream = harness.oracle_encode(data, bs)
        assert len(stream) > 4 * 32768
        rc, got = emu.decode(stream)
        assert (rc, got) == (0, data), bs
    # false headers inside a payload: a foreign block (identity-like tree over raw bytes) whose
    # payload carries the signature bytes of a plausible header at several offsets
    from cases import hdr
    leafs = []
    def full(depth, prefix):
        if depth == 8:
            return [prefix & 0xff, -1, -1]
        return [300 + depth] + full(depth + 1, prefix << 1) + full(depth + 1, (prefix << 1) | 1)
    tree = full(0, 0)                                    # 8-bit identity code, binary root
    raw = bytearray(rng.integers(0, 256, 120000, dtype=np.uint8).tobytes())
    fake = hdr(4096, [256 + 3, 1, -1, -1])[:12]          # orig_len, tree_len = 4n+1?, tree[0] = 255+n
    fake = (4096).to_bytes(8, "little") + (13).to_bytes(2, "little") + (258).to_bytes(2, "little")
    for at in (5000, 33000, 40007, 70001, 99990):
        raw[at:at + len(fake)] = fake
    stream = hdr(len(raw), tree) + bytes(raw) + harness.oracle_encode(data[:9000], 3000)
    rc_o, out_o, _ = harness.oracle_decode(stream)
    rc, got = emu.decode(stream)
    assert (rc, got) == (rc_o, out_o) and rc == 0


def test_emu_encode_pass_pipeline(emu, harness, monkeypatch):
    """The pipelined encode (passes over workspace slots and side streams, huf_b200.cu
    encode_enqueue) with thresholds shrunk so that a small input runs 8 passes over 3 slots:
    slot reuse, the chained offset scan, a ragged last pass.  Byte-equal to the oracle and to the
    single-stream order."""
    monkeypatch.setenv("HUF_B200_ENC_PIPE_MIN", "1")
    monkeypatch.setenv("HUF_B200_ENC_PIPE_PASS", "1")
    monkeypatch.setenv("HUF_B200_ENC_SLOTS", "3")
    data = datagen.zipf(1000 * 61 + 17, 200, seed=77)
    bs = 1000   # 62 blocks: 8 passes of 8 blocks, the last of 6 (ragged block at the end)
    want = harness.oracle_encode(data, bs)
    for no_overlap in (False, True):
        codec = DeviceCodec(emu)
        try:
            emu.check(emu.dll.huf_b200_ctx_set_option(codec.ctx, 4, int(no_overlap)), "set_option")
            src = C.create_string_buffer(data, len(data))
            cap = codec.encode_bound(len(data), bs)
            dst = C.create_string_buffer(cap)
            codec.encode_async(C.addressof(src), len(data), bs, C.addressof(dst), cap)
            n = codec.encode_finish()
            assert dst.raw[:n] == want, no_overlap
            launches = codec.launches()
            assert launches == (8 * 7 if not no_overlap else 7), launches   # 7 kernels per pass
            ptr, nb = codec.block_offsets()
            offs = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), (nb + 1,)).copy()
            assert nb == 62 and offs[0] == 0 and offs[-1] == n
        finally:
            codec.close()


def test_emu_page_locked_memstreams_are_used_in_place(emu, harness, monkeypatch):
    """Memory streams that live through several codec calls are page-locked by the library
    (streams.c huf__memstream_borrow) and the host lanes then copy straight from / into their
    buffers (huf_b200.cu: src_direct / sink_direct).  Same bytes as the bounce path; a sink that
    has to grow in the middle of a call falls back and is locked again later; closing releases."""
    import libhuffman_b200.capi as capi
    monkeypatch.setenv("HUF_B200_PIN_AFTER", "2")
    monkeypatch.setenv("HUF_B200_PIN_MIN_BYTES", str(64 << 10))   # (the floor is 32 MiB outside tests)
    monkeypatch.setenv("HUF_B200_SPAN_BYTES", str(64 << 10))
    n = (300 << 10) + 777                    # 5 spans
    data = datagen.zipf(n, 200, seed=9)
    bs = 8192
    want = harness.oracle_encode(data, bs)
    src, mid, dst = emu.memstream(n), emu.memstream(len(want) + 4096), emu.memstream(200 << 10)
    direct = [emu.dll.huf_b200_direct_copy_count()]
    try:
        for it in range(4):
            for s_ in (src, mid, dst):
                emu.dll.huf_memrewind(s_.rw)
            src.write(data)
            cfg = capi.Config(length=n, blocksize=bs, reader=src.rw, writer=mid.rw)
            assert emu.dll.huf_encode(C.byref(cfg)) == 0
            assert C.string_at(mid.buf, len(mid)) == want, it
            cfg = capi.Config(length=len(mid), reader=mid.rw, writer=dst.rw)
            assert emu.dll.huf_decode(C.byref(cfg)) == 0
            # (dst starts too small: it grows inside the first call, is locked from its second use
            # on, and the calls after that copy straight into it)
            assert len(dst) == n and C.string_at(dst.buf, n) == data, it
            direct.append(emu.dll.huf_b200_direct_copy_count())
        # first call: everything through the bounce buffers; from the second use of a stream on
        # its buffer is used in place (mid is used twice per round: locked within the first)
        steps = np.diff(direct)
        assert steps[0] > 0 and steps[1] > steps[0] and steps[3] == steps[2] >= 3 * 5 + 1, steps
    finally:
        for s_ in (src, mid, dst):
            s_.close()


@pytest.mark.parametrize("blocksize,extra", [(1 << 18, 4097), (1 << 20, 70001)])
def test_emu_general_packing_lane_whole_rows(emu, harness, blocksize, extra):
    """k_pack_wide packs whole rows of 512 symbols the way the fast lane does (one put per code
    word of up to 26 bits: Fibonacci counts in 256 KiB give 25-bit code words; two puts per code
    word of up to 56 bits: 1 MiB gives 28 bits) and hands the ragged end of a segment to its
    general loop; the second block is short and ends inside a row."""
    data = datagen.fibonacci_block(blocksize, seed=6) + datagen.fibonacci_block(extra, seed=7)
    rc, got = emu.encode(data, blocksize)
    assert rc == 0
    assert got == harness.oracle_encode(data, blocksize)


def test_emu_block_headers_at_every_alignment(emu, harness):
    """The block header leaves in 16-byte lines assembled from the workspace copy of the tree and
    shifted to the header's byte alignment (emit_block_header): blocks of varying length and
    alphabet put headers of 21 ... 2060 bytes at every residue modulo 16."""
    rng = np.random.default_rng(11)
    parts, bs = [], 2500
    for i in range(48):
        nsym = int(rng.integers(2, 257))
        parts.append(rng.integers(0, nsym, size=bs, dtype=np.uint8).tobytes())
    data = b"".join(parts) + b"ab" * 7
    want = harness.oracle_encode(data, bs)
    # block sizes from per-block encodes (every block is coded on its own)
    residues, off = set(), 0
    for i in range(0, len(data), bs):
        residues.add(off % 16)
        off += len(harness.oracle_encode(data[i:i + bs], bs))
    assert off == len(want) and len(residues) == 16
    rc, got = emu.encode(data, bs)
    assert rc == 0 and got == want


def _random_case(rng):
    """A random input: alphabet size, skew (uniform ... a few dominant symbols ... Fibonacci-like
    counts), length and block size all drawn, so that tie-breaks in the merge, deep and flat
    trees, ragged last blocks and one-symbol blocks all turn up."""
    n = int(rng.integers(1, 40000))
    nsym = int(rng.integers(1, 257))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        data = rng.integers(0, nsym, size=n, dtype=np.uint8)
    elif kind == 1:
        p = 1.0 / np.arange(1, nsym + 1) ** float(rng.uniform(0.5, 2.5))
        data = rng.choice(nsym, size=n, p=p / p.sum()).astype(np.uint8)
    elif kind == 2:
        data = np.minimum(rng.geometric(float(rng.uniform(0.2, 0.7)), size=n) - 1, nsym - 1).astype(np.uint8)
    else:  # runs of equal counts: many ties for the merge order
        reps = int(rng.integers(1, 6))
        data = np.repeat(rng.permutation(nsym).astype(np.uint8), reps)
        data = np.resize(data, n)
        rng.shuffle(data)
    perm = rng.permutation(256).astype(np.uint8)  # symbols are not ranks
    bs = int(rng.choice([0, 300, 1000, 4096, 16384, 65536]))
    out = perm[data].tobytes()
    return (out[:20 * bs] if bs else out), bs  # (at most 20 blocks: the emulator runs a CTA at a time)


@pytest.mark.parametrize("seed", range(12))
def test_emu_random_inputs_bit_exact(emu, harness, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(8):
        data, bs = _random_case(rng)
        want = harness.oracle_encode(data, bs)
        rc, got = emu.encode(data, bs)
        assert rc == 0 and got == want, (seed, len(data), bs)
        # device entry points with the lenient 1025-element mode, so 256-symbol blocks decode too
        rc, back, _ = _lanes(emu, want, len(data), accept_1025=True)
        assert rc == 0 and back == data, (seed, len(data), bs)
