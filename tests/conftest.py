"""Shared fixtures.  `-m "not gpu"` runs on the CPU box: oracle vs golden vectors, host C layer,
C-ABI symbol checks and the kernel-logic emulation.  `-m gpu` runs the parity tests proper on a
B200 through the C-ABI of the real CUDA library."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "vectors.json").read_text())


@pytest.fixture(scope="session")
def harness():
    from oracle import harness as h
    h.build()
    return h


@pytest.fixture(scope="session")
def product_path():
    """Path of the real CUDA library (built in-tree by nvcc; cross-compiles without a GPU)."""
    from libhuffman_b200 import LIB_PATH, build
    if not LIB_PATH.exists():
        build.build()
    return LIB_PATH


@pytest.fixture(scope="session")
def lib(product_path):
    """The product library on a GPU box."""
    import libhuffman_b200
    return libhuffman_b200.load()


@pytest.fixture(scope="session")
def emu():
    """Kernel-logic emulation of the library (tests/emu, g++ only).  Test infrastructure."""
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    from libhuffman_b200.capi import B200Lib
    return B200Lib(build_emu.build())
