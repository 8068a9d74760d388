"""Host-side boundary checks that need no GPU: the CUDA library loads, exports every symbol the
headers declare, keeps the reference's struct layouts and stream semantics, and fails loudly
(HUF_ERROR_FATAL, no CPU fallback) when the codec is called without a B200.

The API-object cases mirror the reference's unit tests (test/io_test.c, test/histogram_test.c,
test/tree_test.c, test/symbol_test.c) call for call."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest

from libhuffman_b200.capi import Config, HuffmanCLib, ReadWriter

ROOT = Path(__file__).resolve().parents[1]

REFERENCE_EXPORTS = """huf_error_string huf_memopen huf_memlen huf_memcap huf_memrewind huf_memclose
huf_fdopen huf_fdclose huf_config_init huf_config_free huf_decoder_init huf_decoder_free huf_decode
huf_encoder_init huf_encoder_free huf_encode huf_bit_write huf_bit_read_writer_reset
huf_bufio_read_writer_init huf_bufio_read_writer_free huf_bufio_read_writer_flush huf_bufio_write
huf_bufio_read huf_bufio_read_uint8 huf_bufio_write_uint8 huf_histogram_init huf_histogram_free
huf_histogram_reset huf_histogram_populate huf_malloc huf_symbol_mapping_element_init
huf_symbol_mapping_element_free huf_symbol_mapping_init huf_symbol_mapping_free
huf_symbol_mapping_insert huf_symbol_mapping_get huf_symbol_mapping_reset huf_node_to_string
huf_tree_init huf_tree_free huf_tree_reset huf_tree_deserialize huf_tree_serialize
huf_tree_from_histogram""".split()


@pytest.fixture(scope="module")
def host(product_path):
    return HuffmanCLib(product_path)


def _declared(header: Path) -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    names = set(re.findall(r"\b(huf_[a-z0-9_]+)\s*\(", text))
    return sorted(n for n in names if not n.endswith("_t"))  # drop `huf_error_t (*fn)(...)`


def test_exports_every_declared_symbol(host):
    names = _declared(ROOT / "include" / "huffman.h") + _declared(ROOT / "include" / "huffman" / "b200.h")
    assert len(names) >= 44 + 15
    for name in names:
        assert hasattr(host.dll, name), f"{name} declared in include/ but not exported"


def test_exports_the_references_44_functions(host):
    # the cffi cdef references every one of these (setup_ffi.py:8-23,59-66)
    assert len(REFERENCE_EXPORTS) == 44
    declared = set(_declared(ROOT / "include" / "huffman.h"))
    for name in REFERENCE_EXPORTS:
        assert name in declared, name
        assert hasattr(host.dll, name), name


def test_compat_headers_exist():
    for h in "bufio common config decoder encoder errors histogram io malloc symbol sys tree".split():
        assert (ROOT / "include" / "huffman" / f"{h}.h").exists()


def test_struct_layouts_match_reference_lp64():
    # huf_config_t 48 bytes, huf_read_writer_t 24 bytes
    assert C.sizeof(Config) == 48 and C.sizeof(ReadWriter) == 24
    assert [getattr(Config, f).offset for f, _ in Config._fields_] == [0, 8, 16, 24, 32, 40]
    assert [getattr(ReadWriter, f).offset for f, _ in ReadWriter._fields_] == [0, 8, 16]


def test_error_strings(host):
    want = {0: b"Success", 3: b"Failed on read/write operation", 4: b"Fatal error",
            5: b"Block is corrupted, Huffman tree has impossible size", 7: b"Unknown error", -1: b"Unknown error"}
    for code, text in want.items():
        assert host.dll.huf_error_string(code) == text


# ---- reference test/io_test.c ------------------------------------------------------------------

def test_membuf_write_len_and_buffer_survives_close(host):
    m = host.memstream(256)
    assert len(m) == 0
    m.write(b"membuf test")
    assert len(m) == 11
    buf = m.buf
    host.dll.huf_memclose(C.byref(m.rw))
    assert not m.rw                                  # set to NULL
    assert C.string_at(buf, 11) == b"membuf test"     # data buffer is the caller's
    host.libc.free(buf)


def test_membuf_realloc_2_to_16(host):
    m = host.memstream(2)
    m.write(b"01")
    cap = C.c_size_t()
    host.dll.huf_memcap(m.rw, C.byref(cap))
    assert cap.value == 2
    prev = m.buf.value
    m.write(b"23456789")
    host.dll.huf_memcap(m.rw, C.byref(cap))
    assert cap.value == 16 and m.buf.value != prev
    assert m.getvalue() == b"0123456789"
    m.close()


def test_membuf_read_clamps_and_eof(host):
    m = host.memstream(8)
    m.write(b"abcd")
    assert m.read(8) == b"abcd"
    assert m.read(8) == b""
    m.rewind()
    assert len(m) == 0
    m.close()


def test_membuf_growth_never_overflows(host):
    # Q8: cap=len=10, count=15 -> the reference allocates 20 < 25
    m = host.memstream(10)
    m.write(b"x" * 10)
    m.write(b"y" * 15)
    assert m.getvalue() == b"x" * 10 + b"y" * 15
    m.close()


def test_fd_stream_roundtrip(host, tmp_path):
    import os
    host.dll.huf_fdopen.argtypes = [C.POINTER(C.POINTER(ReadWriter)), C.c_int]
    host.dll.huf_fdclose.argtypes = [C.POINTER(C.POINTER(ReadWriter))]
    path = tmp_path / "f.bin"
    fd = os.open(path, os.O_RDWR | os.O_CREAT)
    rw = C.POINTER(ReadWriter)()
    assert host.dll.huf_fdopen(C.byref(rw), fd) == 0
    data = b"hello fd stream"
    assert rw.contents.write(rw.contents.stream, C.cast(C.c_char_p(data), C.c_void_p), len(data)) == 0
    os.lseek(fd, 0, os.SEEK_SET)
    out = C.create_string_buffer(64)
    n = C.c_size_t(64)
    assert rw.contents.read(rw.contents.stream, C.cast(out, C.c_void_p), C.byref(n)) == 0
    assert out.raw[: n.value] == data
    n = C.c_size_t(64)
    assert rw.contents.read(rw.contents.stream, C.cast(out, C.c_void_p), C.byref(n)) == 0 and n.value == 0
    assert host.dll.huf_fdclose(C.byref(rw)) == 0 and not rw
    os.close(fd)


# ---- reference test/histogram_test.c -----------------------------------------------------------

class Histogram(C.Structure):
    _fields_ = [("frequencies", C.POINTER(C.c_uint64)), ("iota", C.c_size_t), ("length", C.c_size_t),
                ("start", C.c_size_t)]


def test_histogram_object(host):
    d = host.dll
    d.huf_histogram_init.argtypes = [C.POINTER(C.POINTER(Histogram)), C.c_size_t, C.c_size_t]
    d.huf_histogram_populate.argtypes = [C.POINTER(Histogram), C.c_void_p, C.c_size_t]
    d.huf_histogram_reset.argtypes = [C.POINTER(Histogram)]
    d.huf_histogram_free.argtypes = [C.POINTER(C.POINTER(Histogram))]
    h = C.POINTER(Histogram)()
    assert d.huf_histogram_init(C.byref(h), 4, 10) == 0
    assert (h.contents.iota, h.contents.length, h.contents.start) == (4, 10, 2**64 - 1)
    a1 = (C.c_uint32 * 10)(*range(10))
    d.huf_histogram_populate(h, a1, C.sizeof(a1))
    assert h.contents.start == 0 and [h.contents.frequencies[i] for i in range(10)] == [1] * 10
    a2 = (C.c_uint32 * 8)(0, 0, 1, 1, 8, 8, 8, 8)
    d.huf_histogram_populate(h, a2, C.sizeof(a2))
    assert [h.contents.frequencies[i] for i in range(10)] == [3, 3, 1, 1, 1, 1, 1, 1, 5, 1]
    d.huf_histogram_reset(h)
    assert h.contents.start == 2**64 - 1 and h.contents.frequencies[8] == 0
    a3 = (C.c_uint32 * 3)(7, 5, 9)
    d.huf_histogram_populate(h, a3, C.sizeof(a3))
    assert h.contents.start == 5
    assert d.huf_histogram_free(C.byref(h)) == 0 and not h
    assert d.huf_histogram_init(None, 1, 1) == 2  # HUF_ERROR_INVALID_ARGUMENT


# ---- reference test/tree_test.c ------------------------------------------------------------------

class Node(C.Structure):
    pass


Node._fields_ = [("index", C.c_int16), ("parent", C.POINTER(Node)), ("left", C.POINTER(Node)),
                 ("right", C.POINTER(Node))]


class Tree(C.Structure):
    _fields_ = [("leaves", C.POINTER(C.POINTER(Node))), ("root", C.POINTER(Node))]


def test_tree_object_unary_root_and_serialisation(host, harness):
    d = host.dll
    d.huf_histogram_init.argtypes = [C.POINTER(C.POINTER(Histogram)), C.c_size_t, C.c_size_t]
    d.huf_histogram_populate.argtypes = [C.POINTER(Histogram), C.c_void_p, C.c_size_t]
    d.huf_histogram_free.argtypes = [C.POINTER(C.POINTER(Histogram))]
    d.huf_tree_init.argtypes = [C.POINTER(C.POINTER(Tree))]
    d.huf_tree_free.argtypes = [C.POINTER(C.POINTER(Tree))]
    d.huf_tree_reset.argtypes = [C.POINTER(Tree)]
    d.huf_tree_from_histogram.argtypes = [C.POINTER(Tree), C.POINTER(Histogram)]
    d.huf_tree_serialize.argtypes = [C.POINTER(Tree), C.POINTER(C.c_int16), C.POINTER(C.c_size_t)]
    d.huf_tree_deserialize.argtypes = [C.POINTER(Tree), C.POINTER(C.c_int16), C.c_size_t]
    d.huf_node_to_string.argtypes = [C.POINTER(Node), C.c_char_p, C.POINTER(C.c_size_t)]

    h = C.POINTER(Histogram)()
    t = C.POINTER(Tree)()
    assert d.huf_histogram_init(C.byref(h), 1, 512) == 0 and d.huf_tree_init(C.byref(t)) == 0
    arr = (C.c_uint8 * 4)(3, 3, 3, 3)
    d.huf_histogram_populate(h, arr, 4)
    assert d.huf_tree_from_histogram(t, h) == 0
    leaf = t.contents.leaves[3]
    assert leaf and leaf.contents.index == 3
    root = t.contents.root
    assert root and root.contents.index == 256
    assert C.addressof(root.contents.left.contents) == C.addressof(leaf.contents)
    assert not root.contents.right
    d.huf_tree_reset(t)
    assert not t.contents.root

    # a bigger histogram: serialisation and code strings equal the oracle's
    data = b"abracadabra alakazam"
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    d.huf_histogram_free(C.byref(h))
    d.huf_histogram_init(C.byref(h), 1, 512)
    d.huf_histogram_populate(h, buf, len(data))
    assert d.huf_tree_from_histogram(t, h) == 0
    out = (C.c_int16 * 1030)()
    n = C.c_size_t()
    assert d.huf_tree_serialize(t, out, C.byref(n)) == 0
    freq = [data.count(bytes([s])) for s in range(256)]
    lens, codes, tree = harness.oracle_codebook(freq)
    assert list(out)[: n.value] == tree
    for s in set(data):
        sbuf = C.create_string_buffer(64)
        ln = C.c_size_t(64)
        d.huf_node_to_string(t.contents.leaves[s], sbuf, C.byref(ln))
        assert sbuf.raw[: ln.value].decode()[::-1] == codes[s]   # leaf->root chars
    # deserialise what was serialised and dump again
    t2 = C.POINTER(Tree)()
    d.huf_tree_init(C.byref(t2))
    assert d.huf_tree_deserialize(t2, out, n.value) == 0
    out2 = (C.c_int16 * 1030)()
    n2 = C.c_size_t()
    d.huf_tree_serialize(t2, out2, C.byref(n2))
    assert list(out2)[: n2.value] == tree
    d.huf_tree_free(C.byref(t2))
    d.huf_tree_free(C.byref(t))
    d.huf_histogram_free(C.byref(h))
    assert not t and not h


# ---- reference test/symbol_test.c ----------------------------------------------------------------

class SymElem(C.Structure):
    _fields_ = [("length", C.c_size_t), ("coding", C.c_char_p)]


class SymMap(C.Structure):
    _fields_ = [("length", C.c_size_t), ("symbols", C.POINTER(C.POINTER(SymElem)))]


def test_symbol_mapping_object(host):
    d = host.dll
    d.huf_symbol_mapping_element_init.argtypes = [C.POINTER(C.POINTER(SymElem)), C.c_char_p, C.c_size_t]
    d.huf_symbol_mapping_init.argtypes = [C.POINTER(C.POINTER(SymMap)), C.c_size_t]
    d.huf_symbol_mapping_insert.argtypes = [C.POINTER(SymMap), C.c_size_t, C.POINTER(SymElem)]
    d.huf_symbol_mapping_get.argtypes = [C.POINTER(SymMap), C.c_size_t, C.POINTER(C.POINTER(SymElem))]
    d.huf_symbol_mapping_reset.argtypes = [C.POINTER(SymMap)]
    d.huf_symbol_mapping_free.argtypes = [C.POINTER(C.POINTER(SymMap))]
    e = C.POINTER(SymElem)()
    assert d.huf_symbol_mapping_element_init(C.byref(e), b"0101XX", 4) == 0
    assert e.contents.length == 4 and e.contents.coding == b"0101"   # copies `length` bytes + NUL
    m = C.POINTER(SymMap)()
    assert d.huf_symbol_mapping_init(C.byref(m), 256) == 0 and m.contents.length == 256
    assert d.huf_symbol_mapping_insert(m, 65, e) == 0
    e2 = C.POINTER(SymElem)()
    d.huf_symbol_mapping_element_init(C.byref(e2), b"11", 2)
    assert d.huf_symbol_mapping_insert(m, 65, e2) == 0                # previous occupant is freed
    got = C.POINTER(SymElem)()
    assert d.huf_symbol_mapping_get(m, 65, C.byref(got)) == 0 and got.contents.coding == b"11"
    assert d.huf_symbol_mapping_get(m, 256, C.byref(got)) == 2        # out of range
    assert d.huf_symbol_mapping_reset(m) == 0
    d.huf_symbol_mapping_get(m, 65, C.byref(got))
    assert not got
    assert d.huf_symbol_mapping_free(C.byref(m)) == 0 and not m


# ---- bit writer + bufio ------------------------------------------------------------------------------

class BitRW(C.Structure):
    _fields_ = [("bits", C.c_uint8), ("offset", C.c_uint8)]


class Bufio(C.Structure):
    _fields_ = [("bytes", C.POINTER(C.c_uint8)), ("offset", C.c_size_t), ("capacity", C.c_size_t),
                ("length", C.c_size_t), ("have_been_processed", C.c_uint64), ("read_writer", C.POINTER(ReadWriter))]


def test_bit_writer_is_msb_first(host):
    d = host.dll
    d.huf_bit_write.argtypes = [C.POINTER(BitRW), C.c_uint8]
    d.huf_bit_write.restype = None
    d.huf_bit_read_writer_reset.argtypes = [C.POINTER(BitRW)]
    d.huf_bit_read_writer_reset.restype = None
    b = BitRW()
    d.huf_bit_read_writer_reset(C.byref(b))
    assert (b.bits, b.offset) == (0, 8)
    for bit in (1, 0, 1, 1):
        d.huf_bit_write(C.byref(b), bit)
    assert (b.bits, b.offset) == (0b10110000, 4)


def test_bufio_buffered_and_passthrough(host):
    d = host.dll
    d.huf_bufio_read_writer_init.argtypes = [C.POINTER(C.POINTER(Bufio)), C.POINTER(ReadWriter), C.c_size_t]
    d.huf_bufio_read_writer_free.argtypes = [C.POINTER(C.POINTER(Bufio))]
    d.huf_bufio_read_writer_flush.argtypes = [C.POINTER(Bufio)]
    d.huf_bufio_write.argtypes = [C.POINTER(Bufio), C.c_char_p, C.c_size_t]
    d.huf_bufio_read.argtypes = [C.POINTER(Bufio), C.c_char_p, C.c_size_t]
    for cap in (0, 4, 64):
        m = host.memstream(16)
        w = C.POINTER(Bufio)()
        assert d.huf_bufio_read_writer_init(C.byref(w), m.rw, cap) == 0
        for chunk in (b"ab", b"cdefg", b"h", b"ijklmnopqrstuvwxyz"):
            assert d.huf_bufio_write(w, chunk, len(chunk)) == 0
        assert d.huf_bufio_read_writer_flush(w) == 0
        assert m.getvalue() == b"abcdefghijklmnopqrstuvwxyz"
        r = C.POINTER(Bufio)()
        d.huf_bufio_read_writer_init(C.byref(r), m.rw, cap)
        out = C.create_string_buffer(32)
        assert d.huf_bufio_read(r, out, 3) == 0 and out.raw[:3] == b"abc"
        assert d.huf_bufio_read(r, out, 20) == 0 and out.raw[:20] == b"defghijklmnopqrstuvw"
        assert r.contents.have_been_processed == 23
        assert d.huf_bufio_read(r, out, 10) == 3      # short read -> HUF_ERROR_READ_WRITE
        d.huf_bufio_read_writer_free(C.byref(r))
        d.huf_bufio_read_writer_free(C.byref(w))
        m.close()


# ---- codec entry points without a GPU ------------------------------------------------------------------

def test_codec_argument_checks_and_empty_input(host):
    assert host.dll.huf_encode(None) == 2 and host.dll.huf_decode(None) == 2   # Q5 fix
    cfg = Config(length=5)
    assert host.dll.huf_encode(C.byref(cfg)) == 2                              # NULL reader/writer
    # length == 0 succeeds without output and without touching the GPU (Q12, decode_test.c:22-36)
    rc, out = host.encode(b"", 0)
    assert (rc, out) == (0, b"")
    rc, out = host.decode(b"", 0)
    assert (rc, out) == (0, b"")


def test_codec_fails_loudly_without_gpu(host, product_path):
    from libhuffman_b200.capi import B200Lib
    lib = B200Lib(product_path)
    if lib.dll.huf_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    rc, out = host.encode(b"some data", 4)
    assert rc == 4 and out == b""          # HUF_ERROR_FATAL: no CPU fallback
    rc, out = host.decode(bytes(21), 21)
    assert rc == 4 and out == b""
    ctx = C.c_void_p()
    assert lib.dll.huf_b200_ctx_create(C.byref(ctx), -1) == 4
