"""The oracle is pinned before it is trusted: against the reference's golden vectors
(tests/golden/vectors.json, generated from the unmodified reference) and, when oracle/_ref is
present, against the compiled reference on randomized inputs."""
from __future__ import annotations

import numpy as np
import pytest

from libhuffman_b200 import datagen


def test_oracle_matches_golden_encode(golden, harness):
    for v in golden["encode"]:
        got = harness.oracle_encode(bytes.fromhex(v["input"]), v["blocksize"])
        assert got == bytes.fromhex(v["stream"]), v["name"]


def test_reference_known_answers(golden, harness):
    by = {v["name"]: v for v in golden["encode"]}
    # reference test/encode_test.c:35 — 1 byte at blocksize 256 encodes to 21 bytes
    s = bytes.fromhex(by["ref_encode_test_single_1"]["stream"])
    assert len(s) == 21
    assert s.hex() == "0100000000000000" "0500" "0001" "3100" "ffffffffffff" "00"
    # "aab" -> tree [257,256,'b',-1,-1,'a',-1,-1,-1], payload 0x50
    s = bytes.fromhex(by["survey_aab"]["stream"])
    assert len(s) == 29 and s[-1] == 0x50
    assert np.frombuffer(s[10:28], dtype="<i2").tolist() == [257, 256, 98, -1, -1, 97, -1, -1, -1]
    assert len(bytes.fromhex(by["survey_abracadabra"]["stream"])) == 57
    assert bytes.fromhex(by["survey_abracadabra"]["stream"])[-5:].hex() == "138d182700"
    s = bytes.fromhex(by["ref_encode_test_digits"]["stream"])
    assert len(s) == 98 and s[-6:].hex() == "10326b1ee540"
    for n, size in ((254, 2330), (255, 2339), (256, 2348)):
        assert len(bytes.fromhex(by[f"distinct_{n}"]["stream"])) == size
    # Q1: 256 distinct symbols serialise 1025 elements, root 511, last element -1
    s = bytes.fromhex(by["distinct_256"]["stream"])
    tree = np.frombuffer(s[10:10 + 2050], dtype="<i2")
    assert int(np.frombuffer(s[8:10], dtype="<i2")[0]) == 1025 and tree[0] == 511 and tree[1024] == -1


def test_oracle_matches_golden_decode(golden, harness):
    for v in golden["decode"]:
        rc, out, _ = harness.oracle_decode(bytes.fromhex(v["stream"]), v["length"])
        assert rc == v["rc"], v["name"]
        if rc == 0:
            assert out == bytes.fromhex(v["output"]), v["name"]


def test_unary_root_codebook(harness):
    # reference test/tree_test.c:12-35: {3,3,3,3} -> root 256 with only a left child
    freq = [0] * 256
    freq[3] = 4
    lens, codes, tree = harness.oracle_codebook(freq)
    assert tree == [256, 3, -1, -1, -1] and codes[3] == "0" and lens[3] == 1


def test_oracle_roundtrip_and_lenient_mode(harness):
    data = datagen.zipf(50000, 256, seed=11)
    stream = harness.oracle_encode(data, 8192)
    rc, out, used = harness.oracle_decode(stream, accept_1025=True)
    assert rc == 0 and out == data and used == len(stream)
    # strict mode mirrors the reference: a 1025-element tree is BTREE_OVERFLOW (Q2)
    full = harness.oracle_encode(bytes(range(256)), 0)
    assert harness.oracle_decode(full)[0] == 5
    assert harness.oracle_decode(full, accept_1025=True)[:2] == (0, bytes(range(256)))


def test_oracle_vs_compiled_reference(harness):
    if not harness.reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref = harness.reference()
    rng = np.random.default_rng(5)
    for it in range(120):
        n = int(rng.choice([1, 2, 17, 300, 4096, 20000]))
        nsym = int(rng.choice([1, 2, 3, 9, 64, 255, 256]))
        mode = it % 3
        if mode == 0:
            data = rng.integers(0, nsym, n, dtype=np.uint8).tobytes()
        elif mode == 1:
            data = np.minimum(rng.geometric(0.3, n) - 1, nsym - 1).astype(np.uint8).tobytes()
        else:
            data = (rng.integers(0, nsym, n) // 3).astype(np.uint8).tobytes()  # many ties
        bs = int(rng.choice([0, 1, 13, 256, 4096, 65536]))
        for rb, wb in ((0, 0), (128, 128)):
            rc, stream = ref.encode(data, bs, rb, wb)
            assert rc == 0
            assert stream == harness.oracle_encode(data, bs), (it, n, nsym, bs)
        rc_ref, out_ref = ref.decode(stream)
        rc_o, out_o, _ = harness.oracle_decode(stream)
        assert rc_ref == rc_o, (it, rc_ref, rc_o)
        if rc_ref == 0:
            assert out_ref == out_o == data


def test_oracle_vs_reference_corrupted_streams(harness):
    """Error-code parity of the restatement on bit-flipped / truncated streams."""
    if not harness.reference_available():
        pytest.skip("oracle/_ref not built")
    ref = harness.reference()
    rng = np.random.default_rng(9)
    base = harness.oracle_encode(datagen.english_text(3000, seed=3), 1024)
    for it in range(150):
        s = bytearray(base)
        kind = it % 3
        if kind == 0:
            s = s[: int(rng.integers(1, len(s)))]
        elif kind == 1:
            for _ in range(3):
                s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        else:
            pos = int(rng.integers(0, len(s) - 4))
            s[pos:pos + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        s = bytes(s)
        # the reference dereferences NULL for an absent root (Q4): skip what would crash it
        tl = int.from_bytes(s[8:10], "little", signed=True) if len(s) >= 10 else 1
        if tl == 0 or (len(s) >= 12 and s[10:12] == b"\xff\xff"):
            continue
        rc_o, out_o, _ = harness.oracle_decode(s)
        if _may_crash_reference(s):
            continue
        rc_ref, out_ref = ref.decode(s)
        assert rc_ref == rc_o, (it, kind, rc_ref, rc_o)
        if rc_ref == 0:
            assert out_ref == out_o


def _may_crash_reference(stream: bytes) -> bool:
    """Walk the block headers: any block with an absent root and orig_len > 0 segfaults the
    reference (Q4); the oracle returns BTREE_CORRUPTED there instead."""
    pos = 0
    from oracle import harness
    # cheap approximation: decode with the oracle block by block and look at each tree head
    while pos + 10 <= len(stream):
        ol = int.from_bytes(stream[pos:pos + 8], "little")
        tl = int.from_bytes(stream[pos + 8:pos + 10], "little", signed=True)
        if tl < 0 or tl > 1024:
            return False
        if ol > 0 and (tl == 0 or stream[pos + 10:pos + 12] == b"\xff\xff"):
            return True
        rc, _, used = harness.oracle_decode(stream[pos:], length=1)
        if rc != 0 or used == 0:
            return False
        pos += used
    return False
