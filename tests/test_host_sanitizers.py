"""ASan + UBSan lane for the host C layer (stands in for the reference's valgrind
re-run of every C test, test/CMakeLists.txt:8-26).  csrc/host/*.c is compiled with
-fsanitize=address,undefined and linked with a stub of the CUDA shim (tests/san/shim_stub.c); a
driver of our own and the reference's four unit-level cmocka programs (histogram, tree, symbol,
io -- the ones that do not need the codec) run under it with leak detection on."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HOST = sorted((ROOT / "libhuffman_b200" / "csrc" / "host").glob("*.c"))
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-g", "-O1"]
REF_TESTS = Path("/root/reference/test")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not present")


def _build(tmp_path, name, main_src, extra_inc=()):
    exe = tmp_path / name
    cmd = ["gcc", "-std=gnu99", *SAN, "-Wall", "-I", str(ROOT / "include"), *sum((["-I", str(i)] for i in extra_inc), []),
           *map(str, HOST), str(ROOT / "tests" / "san" / "shim_stub.c"), str(main_src), "-o", str(exe), "-lpthread"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def _run(exe):
    env = {"ASAN_OPTIONS": "detect_leaks=1:abort_on_error=0", "UBSAN_OPTIONS": "print_stacktrace=1", "PATH": "/usr/bin:/bin"}
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "ERROR: AddressSanitizer" not in proc.stderr and "runtime error" not in proc.stderr, proc.stderr
    return proc.stdout


def test_host_layer_driver_under_asan_ubsan(tmp_path):
    out = _run(_build(tmp_path, "host_driver", ROOT / "tests" / "san" / "host_driver.c"))
    assert "host sanitizer driver ok" in out


@pytest.mark.skipif(not REF_TESTS.is_dir(), reason="reference checkout not present")
@pytest.mark.parametrize("prog", ["histogram_test", "tree_test", "symbol_test", "io_test"])
def test_reference_unit_programs_under_asan_ubsan(tmp_path, prog):
    exe = _build(tmp_path, prog, REF_TESTS / f"{prog}.c", extra_inc=(ROOT / "tests" / "cmocka_shim", REF_TESTS))
    out = _run(exe)
    assert "FAILED" not in out
