"""Build libhuffman_b200.so in-tree: host C layer (gcc) + sm_100a CUDA shim (nvcc).

The shared library is the product: it exports the reference's whole C API (include/huffman.h)
plus the device entry points of include/huffman/b200.h.  It is built next to this file so it
travels with the repository snapshot to the GPU box (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libhuffman_b200.so"

HOST_SOURCES = sorted((CSRC / "host").glob("*.c"))
CUDA_SOURCES = [CSRC / "cuda" / "huf_b200.cu"]
CUDA_DEPS = sorted((CSRC / "cuda").glob("*.cuh"))
HEADERS = [ROOT / "include" / "huffman.h", ROOT / "include" / "huffman" / "b200.h",
           CSRC / "host" / "internal.h"]

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CC = os.environ.get("CC") or shutil.which("gcc") or "gcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]
CC_FLAGS = ["-std=gnu99", "-O2", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter",
            "-I", str(ROOT / "include"), "-pthread"]


def _run(cmd: list[str], log: Path | None = None) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if stale) and return the path of libhuffman_b200.so."""
    OBJ.mkdir(exist_ok=True)
    objects: list[Path] = []
    for src in HOST_SOURCES:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src, *HEADERS]):
            if verbose:
                print("cc  ", src.name)
            _run([CC, *CC_FLAGS, "-c", str(src), "-o", str(obj)])
        objects.append(obj)
    for src in CUDA_SOURCES:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src, *CUDA_DEPS, *HEADERS]):
            if verbose:
                print("nvcc", src.name)
            _run([NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)], log=OBJ / (src.stem + ".ptxas.log"))
        objects.append(obj)
    if force or _stale(LIB, objects):
        if verbose:
            print("link", LIB.name)
        _run([NVCC, "-shared", "-o", str(LIB), *map(str, objects),
              "-Xlinker", "-Bsymbolic", "-Xlinker", "-soname=libhuffman_b200.so", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
