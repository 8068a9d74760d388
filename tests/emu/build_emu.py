"""TEST INFRASTRUCTURE ONLY: build the kernel-logic emulation of the library with g++.

Compiles the unmodified host C layer and the unmodified CUDA sources (with
tests/emu/cuda_emu.h force-included in place of the CUDA toolchain) into
tests/emu/_build/libhuffman_b200_emu.so.  Nothing outside tests/ loads this file.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "libhuffman_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libhuffman_b200_emu.so"

CXX = shutil.which("g++") or "g++"
CC = shutil.which("gcc") or "gcc"


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("emu build step failed: " + " ".join(map(str, cmd)))


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    deps = [*CSRC.rglob("*.c"), *CSRC.rglob("*.h"), *CSRC.rglob("*.cu"), *CSRC.rglob("*.cuh"),
            HERE / "cuda_emu.h", ROOT / "include" / "huffman.h", ROOT / "include" / "huffman" / "b200.h"]
    if not force and LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in deps):
        return LIB
    objs = []
    for src in sorted((CSRC / "host").glob("*.c")):
        obj = OUT / (src.stem + ".o")
        _run([CC, "-std=gnu99", "-O1", "-g", "-fPIC", "-I", ROOT / "include", "-c", src, "-o", obj])
        objs.append(obj)
    cu = CSRC / "cuda" / "huf_b200.cu"
    obj = OUT / "huf_b200_emu.o"
    _run([CXX, "-std=c++17", "-O1", "-g", "-fPIC", "-x", "c++", "-include", HERE / "cuda_emu.h",
          "-I", ROOT / "include", "-I", CSRC / "cuda", "-Wno-attributes", "-c", cu, "-o", obj])
    objs.append(obj)
    _run([CXX, "-shared", "-o", LIB, *objs, "-Wl,-Bsymbolic", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
