"""ctypes binding of the libhuffman C API (include/huffman.h) and of the device entry
points (include/huffman/b200.h).

`HuffmanCLib(path)` drives any shared library that exports the reference's C API — the
product library, or (from tests only) the compiled reference itself — through the very calls
the reference's own tests make: huf_memopen, stream->write, huf_encode / huf_decode,
huf_memlen, stream->read (reference test/encode_test.c:12-94).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HUF_ERROR_SUCCESS = 0
HUF_ERROR_MEMORY_ALLOCATION = 1
HUF_ERROR_INVALID_ARGUMENT = 2
HUF_ERROR_READ_WRITE = 3
HUF_ERROR_FATAL = 4
HUF_ERROR_BTREE_OVERFLOW = 5
HUF_ERROR_BTREE_CORRUPTED = 6

ERROR_NAMES = {
    0: "HUF_ERROR_SUCCESS", 1: "HUF_ERROR_MEMORY_ALLOCATION", 2: "HUF_ERROR_INVALID_ARGUMENT",
    3: "HUF_ERROR_READ_WRITE", 4: "HUF_ERROR_FATAL", 5: "HUF_ERROR_BTREE_OVERFLOW",
    6: "HUF_ERROR_BTREE_CORRUPTED",
}

WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
READ_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t))


class ReadWriter(C.Structure):
    """huf_read_writer_t (include/huffman.h; reference include/huffman/io.h:11-21)."""
    _fields_ = [("stream", C.c_void_p), ("write", WRITE_FN), ("read", READ_FN)]


class Config(C.Structure):
    """huf_config_t (include/huffman.h; reference include/huffman/config.h:10-36)."""
    _fields_ = [
        ("length", C.c_uint64),
        ("blocksize", C.c_uint64),
        ("reader_buffer_size", C.c_size_t),
        ("writer_buffer_size", C.c_size_t),
        ("reader", C.POINTER(ReadWriter)),
        ("writer", C.POINTER(ReadWriter)),
    ]


class HuffmanError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"{what}: {ERROR_NAMES.get(code, code)}")
        self.code = code


class MemStream:
    """huf_memopen / huf_memclose wrapper that frees the caller-owned buffer."""

    def __init__(self, lib: "HuffmanCLib", capacity: int = 0):
        self.lib = lib
        self.buf = C.c_void_p()
        self.rw = C.POINTER(ReadWriter)()
        lib.check(lib.dll.huf_memopen(C.byref(self.rw), C.byref(self.buf), capacity), "huf_memopen")

    def write(self, data: bytes) -> None:
        if data:
            rw = self.rw.contents
            self.lib.check(rw.write(rw.stream, C.cast(C.c_char_p(data), C.c_void_p), len(data)), "stream.write")

    def __len__(self) -> int:
        n = C.c_size_t()
        self.lib.check(self.lib.dll.huf_memlen(self.rw, C.byref(n)), "huf_memlen")
        return n.value

    def getvalue(self) -> bytes:
        """All bytes ever written (the backing buffer), independent of the read cursor."""
        return C.string_at(self.buf, len(self)) if len(self) else b""

    def read(self, count: int) -> bytes:
        out = C.create_string_buffer(max(count, 1))
        n = C.c_size_t(count)
        rw = self.rw.contents
        self.lib.check(rw.read(rw.stream, C.cast(out, C.c_void_p), C.byref(n)), "stream.read")
        return out.raw[: n.value]

    def rewind(self) -> None:
        self.lib.check(self.lib.dll.huf_memrewind(self.rw), "huf_memrewind")

    def close(self) -> None:
        if self.rw:
            self.lib.dll.huf_memclose(C.byref(self.rw))
            self.lib.libc.free(self.buf)
            self.buf = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class HuffmanCLib:
    def __init__(self, path: str | Path):
        self.path = str(path)
        self.dll = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        d = self.dll
        d.huf_memopen.argtypes = [C.POINTER(C.POINTER(ReadWriter)), C.POINTER(C.c_void_p), C.c_size_t]
        d.huf_memlen.argtypes = [C.POINTER(ReadWriter), C.POINTER(C.c_size_t)]
        d.huf_memcap.argtypes = [C.POINTER(ReadWriter), C.POINTER(C.c_size_t)]
        d.huf_memrewind.argtypes = [C.POINTER(ReadWriter)]
        d.huf_memclose.argtypes = [C.POINTER(C.POINTER(ReadWriter))]
        d.huf_encode.argtypes = [C.POINTER(Config)]
        d.huf_decode.argtypes = [C.POINTER(Config)]
        d.huf_error_string.restype = C.c_char_p
        d.huf_error_string.argtypes = [C.c_int]
        for name in ("huf_memopen", "huf_memlen", "huf_memcap", "huf_memrewind", "huf_memclose",
                     "huf_encode", "huf_decode"):
            getattr(d, name).restype = C.c_int

    @staticmethod
    def check(code: int, what: str) -> None:
        if code != HUF_ERROR_SUCCESS:
            raise HuffmanError(code, what)

    def memstream(self, capacity: int = 0) -> MemStream:
        return MemStream(self, capacity)

    # -- whole-buffer helpers in the style of reference test/encode_test.c ------------------

    def encode(self, data: bytes, blocksize: int = 0, reader_buffer: int = 0,
               writer_buffer: int = 0, length: int | None = None) -> tuple[int, bytes]:
        """huf_encode over memory streams.  Returns (huf_error_t, stream bytes)."""
        with self.memstream(len(data)) as src, self.memstream(64) as dst:
            src.write(data)
            cfg = Config(length=len(data) if length is None else length, blocksize=blocksize,
                         reader_buffer_size=reader_buffer, writer_buffer_size=writer_buffer,
                         reader=src.rw, writer=dst.rw)
            rc = self.dll.huf_encode(C.byref(cfg))
            return rc, dst.getvalue()

    def decode(self, stream: bytes, length: int | None = None, reader_buffer: int = 0,
               writer_buffer: int = 0) -> tuple[int, bytes]:
        """huf_decode over memory streams.  Returns (huf_error_t, decoded bytes)."""
        with self.memstream(len(stream)) as src, self.memstream(64) as dst:
            src.write(stream)
            cfg = Config(length=len(stream) if length is None else length, blocksize=0,
                         reader_buffer_size=reader_buffer, writer_buffer_size=writer_buffer,
                         reader=src.rw, writer=dst.rw)
            rc = self.dll.huf_decode(C.byref(cfg))
            return rc, dst.getvalue()


class B200Lib(HuffmanCLib):
    """The product library: reference C API + the device entry points of huffman/b200.h."""

    def __init__(self, path: str | Path):
        super().__init__(path)
        d = self.dll
        u64 = C.c_uint64
        vp = C.c_void_p
        d.huf_b200_device_count.restype = C.c_int
        d.huf_b200_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
        d.huf_b200_ctx_destroy.argtypes = [C.POINTER(vp)]
        d.huf_b200_ctx_set_option.argtypes = [vp, C.c_int, C.c_int64]
        d.huf_b200_encode_bound.restype = u64
        d.huf_b200_encode_bound.argtypes = [u64, u64]
        d.huf_b200_block_count.restype = u64
        d.huf_b200_block_count.argtypes = [u64, u64]
        d.huf_b200_encode_async.argtypes = [vp, vp, u64, u64, vp, u64, vp]
        d.huf_b200_encode_finish.argtypes = [vp, C.POINTER(u64)]
        d.huf_b200_encode_block_offsets.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        d.huf_b200_decode_async.argtypes = [vp, vp, u64, u64, vp, u64, vp]
        d.huf_b200_decode_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        d.huf_b200_decode_async_at.argtypes = [vp, vp, u64, u64, u64, vp, u64, vp]
        d.huf_b200_decode_range_async.argtypes = [vp, vp, u64, u64, u64, C.c_int, vp, u64, vp]
        d.huf_b200_decode_range_plan.argtypes = [vp, vp, u64, u64, u64, C.c_int, C.POINTER(u64), C.POINTER(u64), vp]
        d.huf_b200_decode_range_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        d.huf_b200_decode_plan.argtypes = [vp, vp, u64, u64, C.POINTER(u64), C.POINTER(u64), vp]
        d.huf_b200_last_launch_count.restype = u64
        d.huf_b200_last_launch_count.argtypes = [vp]
        d.huf_b200_decode_hint_offsets.argtypes = [vp, vp, u64]
        d.huf_b200_last_slow_blocks.restype = u64
        d.huf_b200_last_slow_blocks.argtypes = [vp]
        d.huf_b200_direct_copy_count.restype = u64
        d.huf_b200_direct_copy_count.argtypes = []
        d.huf_b200_host_register.argtypes = [vp, u64]
        d.huf_b200_host_register.restype = C.c_int
        d.huf_b200_host_unregister.argtypes = [vp]
        d.huf_b200_host_unregister.restype = C.c_int
        d.huf_b200_kernel_times.argtypes = [vp, C.c_char_p, u64]
        d.huf_b200_kernel_times.restype = C.c_int
        d.huf_b200_dev_alloc.argtypes = [C.POINTER(vp), u64]
        d.huf_b200_dev_free.argtypes = [vp]
        d.huf_b200_copy_h2d.argtypes = [vp, vp, u64]
        d.huf_b200_copy_d2h.argtypes = [vp, vp, u64]
        for name in ("huf_b200_ctx_create", "huf_b200_ctx_destroy", "huf_b200_ctx_set_option",
                     "huf_b200_encode_async", "huf_b200_encode_finish", "huf_b200_encode_block_offsets",
                     "huf_b200_decode_async", "huf_b200_decode_finish", "huf_b200_decode_plan",
                     "huf_b200_decode_async_at", "huf_b200_decode_range_async", "huf_b200_decode_range_finish",
                     "huf_b200_decode_range_plan",
                     "huf_b200_dev_alloc", "huf_b200_dev_free", "huf_b200_copy_h2d", "huf_b200_copy_d2h"):
            getattr(d, name).restype = C.c_int


class DeviceCodec:
    """Device-resident encode/decode through the C-ABI shim (huffman/b200.h).

    Pointers are raw device addresses (e.g. torch.Tensor.data_ptr()); `stream` is a
    cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream) or 0 for the context's
    own stream.
    """

    def __init__(self, lib: B200Lib, device: int = -1, accept_1025: bool | None = None):
        self.lib = lib
        self.ctx = C.c_void_p()
        lib.check(lib.dll.huf_b200_ctx_create(C.byref(self.ctx), device), "huf_b200_ctx_create")
        if accept_1025 is not None:
            self.set_accept_1025(accept_1025)

    def set_accept_1025(self, on: bool) -> None:
        self.lib.check(self.lib.dll.huf_b200_ctx_set_option(self.ctx, 1, int(on)), "set_option")

    def close(self) -> None:
        if self.ctx:
            self.lib.dll.huf_b200_ctx_destroy(C.byref(self.ctx))

    def encode_bound(self, length: int, blocksize: int) -> int:
        return self.lib.dll.huf_b200_encode_bound(length, blocksize)

    def encode_async(self, d_in: int, length: int, blocksize: int, d_out: int, out_cap: int,
                     stream: int = 0) -> None:
        self.lib.check(self.lib.dll.huf_b200_encode_async(self.ctx, d_in, length, blocksize, d_out,
                                                          out_cap, stream), "huf_b200_encode_async")

    def encode_finish(self) -> int:
        n = C.c_uint64()
        self.lib.check(self.lib.dll.huf_b200_encode_finish(self.ctx, C.byref(n)), "huf_b200_encode_finish")
        return n.value

    def block_offsets(self) -> tuple[int, int]:
        p = C.c_void_p()
        n = C.c_uint64()
        self.lib.check(self.lib.dll.huf_b200_encode_block_offsets(self.ctx, C.byref(p), C.byref(n)),
                       "huf_b200_encode_block_offsets")
        return p.value or 0, n.value

    def decode_async(self, d_in: int, avail: int, length: int, d_out: int, out_cap: int,
                     stream: int = 0) -> None:
        self.lib.check(self.lib.dll.huf_b200_decode_async(self.ctx, d_in, avail, length, d_out,
                                                          out_cap, stream), "huf_b200_decode_async")

    def decode_hint_offsets(self, d_offsets: int, nblocks: int) -> None:
        """Block index (device u64 array) for the next decode_async: skips the header scan."""
        self.lib.check(self.lib.dll.huf_b200_decode_hint_offsets(self.ctx, d_offsets, nblocks),
                       "huf_b200_decode_hint_offsets")

    def decode_finish(self) -> tuple[int, int, int]:
        """Returns (huf_error_t, decoded bytes, consumed bytes)."""
        n = C.c_uint64()
        used = C.c_uint64()
        rc = self.lib.dll.huf_b200_decode_finish(self.ctx, C.byref(n), C.byref(used))
        return rc, n.value, used.value

    def decode_range_async(self, d_in: int, avail: int, start: int, stop: int, d_out: int, out_cap: int,
                           stream: int = 0, start_is_block: bool | None = None) -> None:
        """Decode the blocks that start in [start, stop) of the stream (multi-GPU decode unit).
        start_is_block: `start` is a known block start (default: only when start == 0)."""
        if start_is_block is None:
            start_is_block = start == 0
        self.lib.check(self.lib.dll.huf_b200_decode_range_async(self.ctx, d_in, avail, start, stop,
                                                                int(start_is_block), d_out, out_cap, stream),
                       "huf_b200_decode_range_async")

    def decode_range_plan(self, d_in: int, avail: int, start: int, stop: int, stream: int = 0,
                          start_is_block: bool | None = None) -> int:
        """Decoded size of the blocks whose headers the scan finds in [start, stop)."""
        if start_is_block is None:
            start_is_block = start == 0
        n = C.c_uint64()
        self.lib.check(self.lib.dll.huf_b200_decode_range_plan(self.ctx, d_in, avail, start, stop,
                                                               int(start_is_block), C.byref(n), None, stream),
                       "huf_b200_decode_range_plan")
        return n.value

    def decode_range_finish(self) -> tuple[int, int, int, int]:
        """Returns (huf_error_t, offset of the first block found or None, chain end, decoded bytes)."""
        first, end, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = self.lib.dll.huf_b200_decode_range_finish(self.ctx, C.byref(first), C.byref(end), C.byref(n))
        return rc, (None if first.value == 2 ** 64 - 1 else first.value), end.value, n.value

    def decode_plan(self, d_in: int, avail: int, length: int, stream: int = 0) -> tuple[int, int]:
        n = C.c_uint64()
        nb = C.c_uint64()
        self.lib.check(self.lib.dll.huf_b200_decode_plan(self.ctx, d_in, avail, length, C.byref(n),
                                                         C.byref(nb), stream), "huf_b200_decode_plan")
        return n.value, nb.value

    def launches(self) -> int:
        return self.lib.dll.huf_b200_last_launch_count(self.ctx)

    def slow_blocks(self) -> int:
        """Candidate blocks of the last decode pass that took the general (slow) lane."""
        return self.lib.dll.huf_b200_last_slow_blocks(self.ctx)

    def set_force_lut_add(self, on: bool) -> None:
        """Tests: use the fast decode kernel instance that does not need an 8 KB aligned table."""
        self.lib.check(self.lib.dll.huf_b200_ctx_set_option(self.ctx, 3, int(on)), "set_option")

    def set_kernel_timing(self, on: bool) -> None:
        self.lib.check(self.lib.dll.huf_b200_ctx_set_option(self.ctx, 2, int(on)), "set_option")

    def kernel_times(self) -> list[tuple[str, float]]:
        """(kernel name, milliseconds) per launch of the last call (needs set_kernel_timing)."""
        buf = C.create_string_buffer(8192)
        self.lib.check(self.lib.dll.huf_b200_kernel_times(self.ctx, buf, 8192), "huf_b200_kernel_times")
        out = []
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            out.append((name, float(ms)))
        return out
