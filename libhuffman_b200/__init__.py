"""B200-native drop-in for libhuffman's block encode/decode path.

The product is `libhuffman_b200.so` (C API of include/huffman.h + device entry points of
include/huffman/b200.h).  This package only locates/builds it and offers a ctypes binding.
There is no CPU implementation: without the compiled CUDA library `load()` raises, and
without a B200 the codec calls return HUF_ERROR_FATAL.
"""
from __future__ import annotations

import os
from pathlib import Path

# The encoder's pass pipeline runs over 16 CUDA streams; the driver's default of 8 hardware work
# queues makes them alias (huf_b200.cu, huf_b200_ctx_create).  The library sets this itself when
# it is loaded; set here as well because the package is normally imported before torch touches
# the device while the library is loaded after.  An explicit setting of the application wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .capi import (B200Lib, Config, DeviceCodec, HuffmanCLib, HuffmanError, MemStream,  # noqa: F401
                   ReadWriter)

LIB_PATH = Path(__file__).resolve().parent / "libhuffman_b200.so"

_lib: B200Lib | None = None


def load() -> B200Lib:
    """Load the in-tree CUDA library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m libhuffman_b200.build` (needs nvcc). "
                "There is no CPU fallback for the codec.")
        _lib = B200Lib(LIB_PATH)
    return _lib


def compress(data: bytes, blocksize: int = 131072) -> bytes:
    """huf_encode through the C API with host buffers (memory streams)."""
    rc, out = load().encode(data, blocksize)
    if rc:
        raise HuffmanError(rc, "huf_encode")
    return out


def decompress(stream: bytes) -> bytes:
    """huf_decode through the C API with host buffers (memory streams)."""
    rc, out = load().decode(stream)
    if rc:
        raise HuffmanError(rc, "huf_decode")
    return out
