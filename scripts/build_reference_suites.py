#!/usr/bin/env python
"""Build the reference's own test suites against the REAL library for the GPU box.

The GPU box has no reference checkout, so what needs it is built here (where /root/reference
is) into oracle/_ref_suites/ -- git-ignored like oracle/_ref/, but shipped by gpurun:

  oracle/_ref_suites/huffmanfile_gpu/huffmanfile/   the reference's Python package, untouched, plus
                                                 its cffi extension `_C` linked against
                                                 libhuffman_b200.so (scripts/build_huffmanfile_ffi.py)
  oracle/_ref_suites/cmocka_gpu/<program>           the reference's six cmocka programs (test/*.c,
                                                 unmodified) linked against libhuffman_b200.so

Nothing of this is product code; tests/test_reference_suites_gpu.py runs it with -m gpu.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = ROOT / "oracle" / "_ref_suites"
PROGRAMS = ["encode_test", "decode_test", "histogram_test", "tree_test", "symbol_test", "io_test"]


def build(verbose: bool = False) -> bool:
    if not (REF / "huffmanfile").is_dir():
        return False
    lib = ROOT / "libhuffman_b200" / "libhuffman_b200.so"
    if not lib.exists():
        raise RuntimeError("build the library first")
    sys.path.insert(0, str(ROOT / "scripts"))
    import build_huffmanfile_ffi

    pkg_out = OUT / "huffmanfile_gpu"
    stamp = pkg_out / ".built_for"
    sig = f"{lib.stat().st_mtime_ns}"
    if not stamp.exists() or stamp.read_text() != sig:
        if pkg_out.exists():
            shutil.rmtree(pkg_out)
        build_huffmanfile_ffi.build(REF / "huffmanfile", pkg_out, lib)
        stamp.write_text(sig)
        if verbose:
            print("built", pkg_out)
    cm_out = OUT / "cmocka_gpu"
    cm_out.mkdir(parents=True, exist_ok=True)
    for prog in PROGRAMS:
        exe = cm_out / prog
        src = REF / "test" / f"{prog}.c"
        if exe.exists() and exe.stat().st_mtime >= max(lib.stat().st_mtime, src.stat().st_mtime):
            continue
        cmd = ["gcc", "-std=gnu99", "-O1", "-I", str(ROOT / "tests" / "cmocka_shim"), "-I", str(ROOT / "include"),
               "-I", str(REF / "test"), str(src), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(proc.stderr)
        if verbose:
            print("built", exe)
    return True


if __name__ == "__main__":
    print(build(verbose=True))
