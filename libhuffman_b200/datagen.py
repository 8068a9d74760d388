"""Deterministic synthetic inputs of the shapes BASELINE.json names.

numpy versions are used by the tests and the CPU baselines; `*_torch` versions generate the
same distributions directly in HBM for the 1 GiB bench (different bit streams, same law).
"""
from __future__ import annotations

import numpy as np

# Relative frequencies of an English-like order-0 source (letters, space, punctuation).
_ENGLISH = {
    " ": 18.3, "e": 10.3, "t": 7.5, "a": 6.5, "o": 6.2, "n": 5.7, "i": 5.7, "s": 5.3, "r": 5.0,
    "h": 5.0, "l": 3.3, "d": 3.3, "u": 2.3, "c": 2.2, "m": 2.0, "f": 2.0, "w": 1.7, "g": 1.6,
    "p": 1.5, "y": 1.4, "b": 1.3, "v": 0.8, "k": 0.6, ",": 0.9, ".": 0.8, "\n": 0.5, "x": 0.14,
    "j": 0.13, "q": 0.08, "z": 0.06, "T": 0.3, "A": 0.25, "I": 0.25, "S": 0.2, "'": 0.2, "-": 0.15,
    "\"": 0.2, ";": 0.05, ":": 0.05, "?": 0.05, "!": 0.03, "0": 0.05, "1": 0.06, "2": 0.04,
    "H": 0.12, "W": 0.1, "B": 0.1, "M": 0.1, "C": 0.1, "E": 0.08, "N": 0.08, "O": 0.08, "(": 0.03,
    ")": 0.03, "3": 0.02, "9": 0.02, "5": 0.02, "L": 0.05, "D": 0.05, "R": 0.05, "P": 0.05,
}


def english_text(n: int, seed: int = 1) -> bytes:
    """Config 1: order-0 sample of an English-like letter/space/punctuation table."""
    rng = np.random.default_rng(seed)
    syms = np.frombuffer("".join(_ENGLISH).encode("latin-1"), dtype=np.uint8)
    p = np.array(list(_ENGLISH.values()), dtype=np.float64)
    p /= p.sum()
    return syms[rng.choice(len(syms), size=n, p=p)].tobytes()


def zipf_cdf(nsym: int = 256, s: float = 1.1) -> np.ndarray:
    w = (np.arange(nsym, dtype=np.float64) + 1.0) ** (-s)
    return np.cumsum(w / w.sum())


def zipf(n: int, nsym: int = 256, s: float = 1.1, seed: int = 2) -> bytes:
    """Config 2: p(k) ~ (k+1)^-s, byte value = k, inverse-CDF sampling."""
    rng = np.random.default_rng(seed)
    cdf = zipf_cdf(nsym, s)
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for lo in range(0, n, step):
        m = min(step, n - lo)
        out[lo:lo + m] = np.minimum(np.searchsorted(cdf, rng.random(m)), nsym - 1).astype(np.uint8)
    return out.tobytes()


def uniform(n: int, nsym: int = 256, seed: int = 3) -> bytes:
    """Config 3(i): incompressible bytes."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, nsym, size=n, dtype=np.uint8).tobytes()


def fibonacci_block(blocksize: int, seed: int = 4) -> bytes:
    """Config 3(ii): one block whose symbol counts follow 1,1,2,3,5,... (deepest possible
    tree for its size), remainder added to the most frequent symbol, shuffled."""
    fib = [1, 1]
    while sum(fib) + fib[-1] + fib[-2] <= blocksize:
        fib.append(fib[-1] + fib[-2])
    counts = fib[:]
    counts[-1] += blocksize - sum(counts)
    block = np.repeat(np.arange(len(counts), dtype=np.uint8), counts)
    np.random.default_rng(seed).shuffle(block)
    return block.tobytes()


def fibonacci(n: int, blocksize: int = 65536, seed: int = 4) -> bytes:
    out = bytearray()
    i = 0
    while len(out) < n:
        out += fibonacci_block(min(blocksize, n - len(out)), seed + i)
        i += 1
    return bytes(out)


def geometric(n: int, seed: int = 4) -> bytes:
    """Config 3(ii) variant: p(k) = 2^-(k+1)."""
    rng = np.random.default_rng(seed)
    return np.minimum(rng.geometric(0.5, size=n) - 1, 255).astype(np.uint8).tobytes()


# ---- device-side generators for the bench (torch is plumbing here) ---------------------------

def zipf_torch(n: int, device, nsym: int = 256, s: float = 1.1, seed: int = 2):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cdf = torch.tensor(zipf_cdf(nsym, s), dtype=torch.float32, device=device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 26
    for lo in range(0, n, step):
        m = min(step, n - lo)
        u = torch.rand(m, generator=g, device=device)
        out[lo:lo + m] = torch.clamp(torch.searchsorted(cdf, u), max=nsym - 1).to(torch.uint8)
    return out


def uniform_torch(n: int, device, nsym: int = 256, seed: int = 3):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, nsym, (n,), generator=g, device=device, dtype=torch.uint8)
