"""Seeded inputs shared by the emulation lane (CPU) and the GPU parity lane."""
from __future__ import annotations

import struct

import numpy as np

from libhuffman_b200 import datagen


def small_cases():
    """(name, data, blocksize): sizes the oracle and the emulator finish in well under a second."""
    rng = np.random.default_rng(21)
    c = [
        ("one_byte", b"1", 256),
        ("one_byte_bs0", b"x", 0),
        ("aab", b"aab", 0),
        ("digits", b"0123456789", 0),
        ("single_symbol_run", b"a" * 1000, 131072),
        ("single_symbol_multi_block", b"z" * 1000, 96),
        ("two_symbols", b"ab" * 300, 0),
        ("bs1", b"hello world", 1),
        ("bs_odd_unaligned", datagen.english_text(3001, seed=5), 333),
        ("bs17", datagen.english_text(400, seed=6), 17),
        ("ties_all_equal", bytes(range(64)) * 4, 0),
        ("ties_pow2", bytes(sum(([s] * (1 << (s % 6)) for s in range(40)), [])), 0),
        ("distinct_255", bytes(range(255)), 0),
        ("distinct_256", bytes(range(256)), 0),
        ("distinct_256_x3_bs300", bytes(range(256)) * 3, 300),
        ("english_20k_bs4096", datagen.english_text(20000, seed=1), 4096),
        ("zipf256_40k_bs16k", datagen.zipf(40000, 256, seed=2), 16384),
        ("zipf255_70k_bs64k", datagen.zipf(70000, 255, seed=2), 65536),
        ("uniform_33k_bs8k", datagen.uniform(33000, 256, seed=3), 8192),
        ("fibonacci_64k", datagen.fibonacci(65536, 65536, seed=4), 65536),
        ("geometric_30k_bs10000", datagen.geometric(30000, seed=4), 10000),
        ("segment_edge_16384", datagen.zipf(16384, 64, seed=8), 16384),
        ("segment_edge_16385", datagen.zipf(16385, 64, seed=8), 16385),
        ("segment_edge_32769", datagen.zipf(32769 + 16, 200, seed=9), 32769),
        ("short_last_block", datagen.zipf(3 * 4096 + 5, 100, seed=10), 4096),
        ("random_many_ties", (rng.integers(0, 200, 9000) // 7).astype(np.uint8).tobytes(), 2048),
        # half the block is a run of one symbol (2-bit code), half is spread over 255 symbols: the
        # average code length oversizes the decoder's optimistic sub-blocks inside the run
        ("run_then_uniform_64k", bytes([7]) * 32768 + datagen.uniform(32768, 255, seed=6), 65536),
    ]
    return c


def hdr(orig_len: int, tree: list[int]) -> bytes:
    return struct.pack("<Qh", orig_len, len(tree)) + struct.pack(f"<{len(tree)}h", *tree)


def foreign_streams():
    """Streams the reference decoder accepts although its encoder never emits them
    (SURVEY.md §5.2 grammar): name, stream."""
    leaf = lambda s: [s, -1, -1]
    # binary root, no unary wrapper: A=0 B=10 C=11
    t1 = [300] + leaf(65) + [301] + leaf(66) + leaf(67)
    s1 = hdr(5, t1) + bytes([0b01011100, 0b10000000])
    # leaf label above 255 is truncated to a byte (321 -> 'A'); trailing elements ignored
    s2 = hdr(3, [256, 321, -1, -1, -1, 9, 9]) + bytes([0])
    # code longer than the 12-bit table: left chain of depth 14
    s3 = hdr(3, list(range(1, 15)) + [88]) + bytes(6)
    # two different foreign blocks back to back, then a normal one is appended by the caller
    return [("binary_root", s1), ("label_321_trailing", s2), ("deep_chain", s3),
            ("foreign_concat", s1 + s2 + s3)]
