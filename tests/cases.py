"""Seeded input families shared by the oracle and GPU parity tests.

Shape follows lz4net's differential ConformanceTests (src/LZ4.Tests/ConformanceTests.cs:21-46: random lengths
exp(U * ln max), every encoder byte-identical, every decoder round-trips) with synthetic content instead of the
Silesia corpus (no network), plus the upstream fuzzer's generator (original/fuzzer.c:153-168).
"""
from __future__ import annotations

import numpy as np

from lz4net_b200 import synth

LOREM = (
    b"Lorem ipsum dolor sit amet, consectetur adipisicing elit, sed do eiusmod tempor incididunt ut "
    b"labore et dolore magna aliqua. Ut enim ad minim veniam, quis nostrud exercitation ullamco "
    b"laboris nisi ut aliquip ex ea commodo consequat. Duis aute irure dolor in reprehenderit in "
    b"voluptate velit esse cillum dolore eu fugiat nulla pariatur. Excepteur sint occaecat cupidatat "
    b"non proident, sunt in culpa qui officia deserunt mollit anim id est laborum."
)
AUTOTEST = LOREM * 5          # src/LZ4/LZ4Codec.cs:175-184 -- the 2 230-byte startup self-test input

MODELS = ("E0", "E50", "E100", "ETEXT", "runs", "lowent", "mixed", "periodic")


def content(model: str, n: int, seed: int) -> np.ndarray:
    """n bytes of the given content model; deterministic in (model, n, seed)."""
    rng = np.random.default_rng([seed, n, MODELS.index(model)])
    if n == 0:
        return np.zeros(0, np.uint8)
    if model in synth.CLASSES:
        return synth.make_blocks(model, 1, n, seed=seed)[0]
    if model == "runs":            # runs of a repeated byte with geometric lengths (RLE / overlapping matches)
        out = np.empty(n, np.uint8); i = 0
        while i < n:
            l = int(rng.geometric(0.02)); out[i:i + l] = rng.integers(0, 256); i += l
        return out
    if model == "lowent":          # 2-bit alphabet: many short, chancy matches, hash collisions
        return rng.integers(0, 4, n, dtype=np.uint8) + 65
    if model == "periodic":        # short period pattern with sparse mutations: offsets < 8, long matches
        p = int(rng.integers(1, 40)); base = rng.integers(0, 256, p, dtype=np.uint8)
        out = np.tile(base, n // p + 1)[:n].copy()
        k = max(1, n // 500); out[rng.integers(0, n, k)] = rng.integers(0, 256, k, dtype=np.uint8)
        return out
    # mixed: concatenated segments of random / copy-from-earlier / zeros
    out = np.empty(n, np.uint8); i = 0
    while i < n:
        l = int(min(n - i, rng.integers(1, 600))); kind = rng.integers(0, 3)
        if kind == 0 or i == 0:
            out[i:i + l] = rng.integers(0, 256, l, dtype=np.uint8)
        elif kind == 1:
            d = int(rng.integers(1, min(i, 65535) + 1))
            for j in range(l):
                out[i + j] = out[i + j - d]
        else:
            out[i:i + l] = 0
        i += l
    return out


def random_lengths(count: int, max_len: int, seed: int):
    """Lengths drawn like Utilities.cs:35-38: exp(U * ln max_len), plus the boundary sizes that matter."""
    rng = np.random.default_rng(seed)
    ls = [int(np.exp(rng.random() * np.log(max_len))) for _ in range(count)]
    return ls


BOUNDARY_LENGTHS = (0, 1, 2, 4, 5, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 27, 28, 29, 31, 32, 33, 63, 64, 65, 255, 256, 257,
                    269, 270, 271, 272, 273, 4095, 4096, 65535, 65536, 65537, 65545, 65546)
