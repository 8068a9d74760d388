"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * `oracle_*`  — our CPU restatement (oracle/huf_oracle.c -> oracle/libhuf_oracle.so)
  * `reference()` — the UNMODIFIED reference compiled by oracle/Makefile into
                    oracle/_ref/libhuffman_ref.so, driven through its own C API

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libhuf_oracle.so"
REF_SO = HERE / "_ref" / "libhuffman_ref.so"

_oracle = None
_ref = None


def build(ref_root: str = "/root/reference") -> None:
    """Compile the oracle and, when the reference tree is present, oracle/_ref."""
    quiet = {"stdout": subprocess.DEVNULL}
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True, **quiet)
    if Path(ref_root, "src").is_dir():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref", f"REF={ref_root}"], check=True, **quiet)


def _lib():
    global _oracle
    if _oracle is None:
        if not ORACLE_SO.exists():
            build()
        o = C.CDLL(str(ORACLE_SO))
        o.huf_oracle_encode_bound.restype = C.c_uint64
        o.huf_oracle_encode_bound.argtypes = [C.c_uint64, C.c_uint64]
        o.huf_oracle_encode.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64,
                                        C.POINTER(C.c_uint64)]
        o.huf_oracle_decode.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        o.huf_oracle_codebook.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint16), C.c_char_p,
                                          C.POINTER(C.c_int16), C.POINTER(C.c_int)]
        _oracle = o
    return _oracle


def oracle_encode(data: bytes, blocksize: int = 0) -> bytes:
    o = _lib()
    cap = o.huf_oracle_encode_bound(len(data), blocksize) + 64
    out = C.create_string_buffer(cap)
    n = C.c_uint64()
    rc = o.huf_oracle_encode(data, len(data), blocksize, out, cap, C.byref(n))
    if rc:
        raise RuntimeError(f"oracle encode failed: {rc}")
    return out.raw[: n.value]


def oracle_decode(stream: bytes, length: int | None = None, out_cap: int | None = None,
                  accept_1025: bool = False) -> tuple[int, bytes, int]:
    """Returns (huf_error_t, decoded bytes, consumed bytes)."""
    o = _lib()
    if out_cap is None:
        out_cap = 8 * len(stream) + 64
    out = C.create_string_buffer(out_cap)
    n = C.c_uint64()
    used = C.c_uint64()
    rc = o.huf_oracle_decode(stream, len(stream), len(stream) if length is None else length, out,
                             out_cap, C.byref(n), C.byref(used), int(accept_1025))
    return rc, out.raw[: n.value], used.value


def oracle_codebook(freq: list[int]) -> tuple[list[int], list[str], list[int]]:
    o = _lib()
    f = (C.c_uint64 * 256)(*freq)
    ln = (C.c_uint16 * 256)()
    codes = C.create_string_buffer(256 * 512)
    tree = (C.c_int16 * 1030)()
    tl = C.c_int()
    o.huf_oracle_codebook(f, ln, codes, tree, C.byref(tl))
    cs = [codes.raw[s * 512: s * 512 + ln[s]].decode() for s in range(256)]
    return list(ln), cs, list(tree)[: tl.value]


def reference_available() -> bool:
    return REF_SO.exists()


def reference():
    """The compiled reference behind the same ctypes driver the product uses."""
    global _ref
    if _ref is None:
        from libhuffman_b200.capi import HuffmanCLib
        _ref = HuffmanCLib(REF_SO)
    return _ref
