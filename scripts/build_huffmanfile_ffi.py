#!/usr/bin/env python
"""Build the reference's cffi extension `huffmanfile._C` against this library.

The reference's setup_ffi.py (reference setup_ffi.py:8-66) scrapes the text between
`#define CFFI_*` / `#undef CFFI_*` fences of the public headers into a cffi cdef and compiles
src/*.c into the extension.  include/huffman.h keeps such a fence around every declaration, so
the same scraping rule works; the only change is that the extension LINKS the B200 library
instead of compiling C sources.  The package `huffmanfile/*.py` is used unchanged: this script
copies it from a reference checkout into the output directory at build time.

    build_huffmanfile_ffi.py --pkg-src /path/to/libhuffman/huffmanfile --out BUILD_DIR [--lib LIB.so]
"""
from __future__ import annotations

import argparse
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def scrape_cdef(header: Path) -> str:
    """Same rule as the reference's make_library_prototypes (setup_ffi.py:8-23)."""
    out, inside = [], False
    for line in header.read_text().splitlines(keepends=True):
        if line.startswith("#define CFFI"):
            inside = True
            continue
        if line.startswith("#undef CFFI"):
            inside = False
        if inside:
            out.append(line)
    return "".join(out)


def build(pkg_src: Path, out: Path, lib: Path) -> Path:
    import cffi

    out.mkdir(parents=True, exist_ok=True)
    pkg = out / "huffmanfile"
    if pkg.exists():
        shutil.rmtree(pkg)
    shutil.copytree(pkg_src, pkg, ignore=shutil.ignore_patterns("_C*", "__pycache__"))
    ffi = cffi.FFI()
    ffi.set_source(
        "huffmanfile._C",
        "#include <huffman.h>",
        include_dirs=[str(ROOT / "include")],
        extra_link_args=[str(lib), f"-Wl,-rpath,{lib.parent}"],
    )
    ffi.cdef(scrape_cdef(ROOT / "include" / "huffman.h"))
    ffi.compile(tmpdir=str(out), verbose=False)
    return pkg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pkg-src", required=True, type=Path, help="the reference's huffmanfile/ directory")
    ap.add_argument("--out", required=True, type=Path)
    ap.add_argument("--lib", type=Path, default=ROOT / "libhuffman_b200" / "libhuffman_b200.so")
    a = ap.parse_args()
    print(build(a.pkg_src, a.out, a.lib.resolve()))


if __name__ == "__main__":
    sys.exit(main())
