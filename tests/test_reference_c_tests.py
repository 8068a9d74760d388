"""the reference's own cmocka unit tests (test/*.c, 15 cases in 6 programs:
encode, decode errors, histogram, tree, symbol, io) compiled UNMODIFIED from the reference
checkout against this library's headers and linked with the library.  cmocka is not installed,
tests/cmocka_shim/cmocka.h stands in for the handful of names the tests use.  In this lane the
library is the kernel-logic emulation build; source compatibility of include/huffman*.h (struct
fields the tests poke into, constants, signatures) is what is being proven, together with the
reference's golden values (length 21, unary root, error codes, memstream growth 2 -> 16).
Skipped where the reference tree is absent (the GPU box)."""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF_TESTS = Path("/root/reference/test")
PROGRAMS = ["encode_test", "decode_test", "histogram_test", "tree_test", "symbol_test", "io_test"]

pytestmark = pytest.mark.skipif(not REF_TESTS.is_dir() or shutil.which("gcc") is None,
                                reason="reference checkout or gcc not present")


@pytest.fixture(scope="module")
def emu_lib():
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    return build_emu.build()


@pytest.mark.parametrize("prog", PROGRAMS)
def test_reference_c_program(prog, emu_lib, tmp_path):
    exe = tmp_path / prog
    cmd = ["gcc", "-std=gnu99", "-O1", "-I", str(ROOT / "tests" / "cmocka_shim"), "-I", str(ROOT / "include"),
           "-I", str(REF_TESTS), str(REF_TESTS / f"{prog}.c"), "-o", str(exe),
           str(emu_lib), f"-Wl,-rpath,{emu_lib.parent}", "-lstdc++", "-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "FAILED" not in run.stdout
