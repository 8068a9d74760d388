"""The reference's own suites against the REAL library on the GPU (BASELINE
config 5): its four Python tests and its HuffmanCompressor streaming path through the unchanged
cffi package, and its six cmocka programs.  Everything reference-side was built on the CPU box by
scripts/build_reference_suites.py into oracle/_ref_suites/ (the GPU box has no reference checkout)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BUILT = ROOT / "oracle" / "_ref_suites"
PKG = BUILT / "huffmanfile_gpu"
PROGRAMS = ["encode_test", "decode_test", "histogram_test", "tree_test", "symbol_test", "io_test"]

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (PKG / "huffmanfile").is_dir(), reason="oracle/_ref_suites not built")]


def test_reference_python_tests_on_gpu():
    proc = subprocess.run([sys.executable, "-m", "pytest", "huffmanfile/huffmanfile_test.py", "-q",
                           "-p", "no:cacheprovider"], cwd=PKG, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "4 passed" in proc.stdout


@pytest.mark.parametrize("prog", PROGRAMS)
def test_reference_c_program_on_gpu(prog):
    exe = BUILT / "cmocka_gpu" / prog
    if not exe.exists():
        pytest.skip("not built")
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "FAILED" not in run.stdout


_STREAM_SCRIPT = r'''
import hashlib, sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "{root}")
import numpy as np
import huffmanfile
from libhuffman_b200 import datagen
total = {mib} << 20
chunk = 64 << 20
data = datagen.zipf(chunk, 255, seed=5)                 # one 64 MiB chunk, fed repeatedly
c = huffmanfile.HuffmanCompressor()                      # default blocksize 131072 (huffmanfile.py:26)
h = hashlib.sha256()
first = None
t0 = time.perf_counter()
n = 0
for at in range(0, total, chunk):
    out = c.compress(data)
    if first is None:
        first = out
    h.update(out)
    n += len(out)
out = c.flush()
h.update(out)
n += len(out)
dt = time.perf_counter() - t0
open("first_chunk.bin", "wb").write(first)
print("RESULT", total, n, dt, h.hexdigest())
d = huffmanfile.HuffmanDecompressor()
assert d.decompress(first) == data
'''


def test_huffman_compressor_stream_on_gpu(harness, tmp_path):
    """HuffmanCompressor.compress() + flush() over 512 MiB in 64 MiB chunks, default block size:
    every call's output equals the oracle's encoding of that chunk (whole blocks per call), the
    reference package's HuffmanDecompressor gives the chunk back."""
    from libhuffman_b200 import datagen
    mib = 512
    env = dict(os.environ)
    proc = subprocess.run([sys.executable, "-c", _STREAM_SCRIPT.format(root=ROOT, mib=mib)], cwd=PKG, env=env,
                          capture_output=True, text=True, timeout=1200)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    total, n, dt = int(line[1]), int(line[2]), float(line[3])
    chunk = datagen.zipf(64 << 20, 255, seed=5)
    want = harness.oracle_encode(chunk, 131072)
    got = (PKG / "first_chunk.bin").read_bytes()
    (PKG / "first_chunk.bin").unlink()
    assert got == want
    assert n == len(want) * (mib // 64)
    if harness.reference_available():
        # and the compiled reference agrees on a prefix (it codes ~13 MB/s: 8 MiB)
        rc, ref_stream = harness.reference().encode(chunk[: 8 << 20], 131072)
        assert rc == 0 and ref_stream == want[: len(ref_stream)]
    print(f"HuffmanCompressor stream: {total / dt / 1e9:.2f} GB/s end to end")
