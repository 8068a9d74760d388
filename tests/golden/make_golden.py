"""Generate tests/golden/vectors.json from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_golden.py

Each encode vector is {name, blocksize, input(hex), stream(hex)} with `stream` produced by the
reference's huf_encode over memory streams; each decode vector is {name, stream(hex), length,
rc, output(hex)} with rc/output produced by the reference's huf_decode.  The first entries
are the reference's own test vectors (test/encode_test.c:12-94, test/decode_test.c:12-81,
huffmanfile/huffmanfile_test.py:15-18) and the survey's known-answer vectors.
"""
from __future__ import annotations

import json
import struct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from libhuffman_b200 import datagen  # noqa: E402
from oracle import harness  # noqa: E402


def main() -> None:
    harness.build()
    ref = harness.reference()
    rng = np.random.default_rng(7)
    enc = []

    def add_enc(name: str, data: bytes, blocksize: int) -> bytes:
        rc, stream = ref.encode(data, blocksize)
        assert rc == 0, (name, rc)
        enc.append({"name": name, "blocksize": blocksize, "input": data.hex(), "stream": stream.hex()})
        return stream

    # reference test vectors + survey known answers
    add_enc("ref_encode_test_single_1", b"1", 256)               # test/encode_test.c:12-45 (21 bytes)
    add_enc("ref_encode_test_digits", b"0123456789", 0)          # test/encode_test.c:48-94
    add_enc("survey_aab", b"aab", 0)
    add_enc("survey_abracadabra", b"abracadabra", 0)
    add_enc("py_test_a1000", b"a" * 1000, 131072)                # huffmanfile_test.py:8-12
    add_enc("py_test_z_incremental_block", b"z" * 10000, 131072)
    lorem = (b"Donec rhoncus quis sapien sit amet molestie. Fusce scelerisque vel augue\n"
             b"nec ullamcorper. Nam rutrum pretium placerat. Aliquam vel tristique lorem,")
    add_enc("py_test_lorem", lorem, 131072)
    # ties, alphabet sizes, ragged block sizes
    add_enc("all_equal_4sym", bytes([3, 1, 2, 0] * 8), 0)
    add_enc("ties_pow2", bytes(sum(([s] * (1 << (s % 5)) for s in range(20)), [])), 0)
    for n in (2, 3, 254, 255, 256):
        add_enc(f"distinct_{n}", bytes(range(n)), 0)
    add_enc("distinct_256_twice_bs300", bytes(range(256)) * 2, 300)
    add_enc("english_8k_bs4096", datagen.english_text(8192, seed=1), 4096)
    add_enc("zipf256_6k_bs1000", datagen.zipf(6000, 256, seed=2), 1000)
    add_enc("zipf255_6k_bs2048", datagen.zipf(6000, 255, seed=2), 2048)
    add_enc("uniform_3k_bs1024", datagen.uniform(3000, 256, seed=3), 1024)
    add_enc("fibonacci_4k", datagen.fibonacci(4096, 4096, seed=4), 4096)
    add_enc("geometric_5k_bs777", datagen.geometric(5000, seed=4), 777)
    add_enc("bs1", b"hello world", 1)
    add_enc("bs_gt_len", b"hello world", 4096)
    add_enc("random_ragged", rng.integers(0, 40, 2500, dtype=np.uint8).tobytes(), 333)

    dec = []

    def add_dec(name: str, stream: bytes, length: int | None = None) -> None:
        n = len(stream) if length is None else length
        rc, out = ref.decode(stream, n)
        dec.append({"name": name, "stream": stream.hex(), "length": n, "rc": rc, "output": out.hex()})

    # test/decode_test.c:12-81
    add_dec("ref_decode_empty", b"", 0)
    add_dec("ref_decode_arbitrary_overflow", bytes([10] * 10))
    add_dec("ref_decode_truncated_tree", bytes([8, 0, 0, 0, 0, 0, 0, 0, 8, 0, 10, 10, 10, 10]))
    add_dec("ref_decode_root_is_leaf", struct.pack("<11h", 8, 0, 0, 0, 3, 0, -1, -1, 1, 2, 3))
    add_dec("py_test_corrupted_header", bytes([8, 0, 0, 0, 0, 0, 0, 0, 2, 0]))  # huffmanfile_test.py:15-18
    # grammar corners: binary root, odd labels, trailing elements, truncation
    hdr = lambda n, tree: struct.pack("<Qh", n, len(tree)) + struct.pack(f"<{len(tree)}h", *tree)
    add_dec("binary_root", hdr(4, [300, 65, -1, -1, 66, -1, -1]) + bytes([0b01100000]))
    add_dec("label_321_is_A", hdr(2, [256, 321, -1, -1, -1]) + bytes([0]))
    add_dec("trailing_elements_ignored", hdr(2, [256, 66, -1, -1, -1, 7, 7, 7]) + bytes([0]))
    add_dec("truncated_tree_absent_children", hdr(3, [256, 67]) + bytes([0]))
    add_dec("bit1_at_unary_root", hdr(3, [256, 67, -1, -1, -1]) + bytes([0b01000000]))
    add_dec("payload_eof", hdr(100, [256, 67, -1, -1, -1]) + bytes([0, 0]))
    add_dec("zero_orig_len_block", hdr(0, [256, 67, -1, -1, -1]) + hdr(3, [256, 68, -1, -1, -1]) + bytes([0]))
    good = bytes.fromhex(enc[1]["stream"])
    add_dec("junk_after_block", good + b"\x01\x02\x03")
    add_dec("length_shorter_than_block", good, 5)
    add_dec("two_calls_concatenated", bytes.fromhex(enc[2]["stream"]) + good)
    add_dec("tree_len_1025_rejected", bytes.fromhex(next(e for e in enc if e["name"] == "distinct_256")["stream"]))
    add_dec("deep_left_chain", hdr(3, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 88]) + bytes([0] * 6))

    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libhuffman_ref.so (unmodified reference, gcc -std=c99 -O2)",
           "encode": enc, "decode": dec}
    path = Path(__file__).with_name("vectors.json")
    path.write_text(json.dumps(out, indent=0))
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(enc)} encode / {len(dec)} decode vectors)")


if __name__ == "__main__":
    main()
