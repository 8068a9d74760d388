"""SURVEY.md §8(f)2 / BASELINE config 5: the reference's cffi package `huffmanfile` must keep
working UNCHANGED against this library.  The package is copied from the reference checkout at
test time (never into the repo), its `_C` extension is built with scripts/build_huffmanfile_ffi.py
-- same CFFI-fence scraping rule as the reference's setup_ffi.py, but linking the library
instead of compiling src/*.c -- and the reference's own four pytest cases are run.  In this lane
the library is the kernel-logic emulation build (no GPU here); the extension binds the very same
44 symbols and struct layouts the CUDA build exports.  Skipped where the reference tree is absent
(the GPU box)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF_PKG = Path("/root/reference/huffmanfile")

pytestmark = pytest.mark.skipif(not REF_PKG.is_dir(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    pytest.importorskip("cffi")
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    sys.path.insert(0, str(ROOT / "scripts"))
    import build_emu
    import build_huffmanfile_ffi

    out = tmp_path_factory.mktemp("huffmanfile_ffi")
    build_huffmanfile_ffi.build(REF_PKG, out, build_emu.build())
    return out


def test_reference_python_tests_pass_unchanged(built):
    proc = subprocess.run([sys.executable, "-m", "pytest", "huffmanfile/huffmanfile_test.py", "-q",
                           "-p", "no:cacheprovider"], cwd=built, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "4 passed" in proc.stdout


def test_streaming_compressor_matches_oracle(built, harness):
    """HuffmanCompressor.compress() in pieces + flush(), HuffmanDecompressor, module-level
    compress/decompress (huffmanfile.py:272-432): bytes equal the oracle's for the same blocking."""
    code = r'''
import sys
sys.path.insert(0, ".")
import huffmanfile
data = bytes((i * 7 + (i >> 5)) % 97 for i in range(200000))
bs = 4096
c = huffmanfile.HuffmanCompressor(blocksize=bs)
out = b"".join(c.compress(data[i:i + 10000]) for i in range(0, len(data), 10000)) + c.flush()
assert huffmanfile.HuffmanDecompressor().decompress(out) == data
assert huffmanfile.decompress(huffmanfile.compress(data, blocksize=bs)) == data
sys.stdout.buffer.write(out)
'''
    proc = subprocess.run([sys.executable, "-c", code], cwd=built, capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    data = bytes((i * 7 + (i >> 5)) % 97 for i in range(200000))
    # the compressor encodes every call's whole blocks, then the remainder as one short block
    assert proc.stdout == harness.oracle_encode(data, 4096)
