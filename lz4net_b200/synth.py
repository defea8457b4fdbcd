"""Deterministic synthetic block generators (host/numpy side).

lz4net's own tests pull the Silesia corpus over HTTP (src/LZ4.Tests/Utilities.cs:13,46-61) or use
System.Random(0) (src/LZ4.Tests/PerformanceTests.cs:136-145); neither is reproducible here, so the entropy
classes of SURVEY.md 8(d) are defined by closed formulas over a counter-based splitmix64 stream.  The CUDA twin
of every class lives in csrc/synth.cu (``lz4b200_synth_fill``) and tests assert both produce the same bytes.

Classes (block of ``n`` bytes, block index ``b``, global ``seed``):
  E0     uniform random bytes                                   (0 % compressible)
  E50    64-byte groups: 32 random bytes then the same 32 again (about 58 % ratio, ~1025 sequences / 64 KiB)
  E100   all zero                                               (one offset-1 match)
  ETEXT  words drawn from an 8-word dictionary, space separated (token-dense worst case)
"""
from __future__ import annotations

import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

CLASSES = ("E0", "E50", "E100", "ETEXT")
CLASS_ID = {c: i for i, c in enumerate(CLASSES)}

# 8 dictionary words; each is emitted followed by one space.  Lengths 2..9 so matches land on every alignment.
DICT = (b"lz", b"net", b"code", b"block", b"stream", b"encoder", b"compress", b"blackwell")


def mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 output function (wrapping uint64 arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def block_seeds(seed: int, first_block: int, n_blocks: int) -> np.ndarray:
    idx = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix64((np.uint64(seed) ^ idx) + GOLD)


def _words(sb: np.ndarray, n_words: int) -> np.ndarray:
    """word k of block with seed sb = mix64(sb + (k+1)*GOLD); shape [n_blocks, n_words]."""
    k = np.arange(1, n_words + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix64(sb[:, None] + k[None, :] * GOLD)


def make_blocks(cls: str, n_blocks: int, block_size: int = 65536, seed: int = 1, first_block: int = 0) -> np.ndarray:
    """Return uint8[n_blocks, block_size]."""
    if cls not in CLASS_ID:
        raise ValueError(f"unknown class {cls!r}")
    n = block_size
    sb = block_seeds(seed, first_block, n_blocks)
    if cls == "E100":
        return np.zeros((n_blocks, n), dtype=np.uint8)
    if cls == "E0":
        w = _words(sb, (n + 7) // 8)
        return np.ascontiguousarray(w.view(np.uint8).reshape(n_blocks, -1)[:, :n])
    if cls == "E50":
        groups = (n + 63) // 64
        w = _words(sb, groups * 4).reshape(n_blocks, groups, 4)
        w = np.concatenate([w, w], axis=2)                      # 32 random bytes, then the same 32 bytes
        return np.ascontiguousarray(w.reshape(n_blocks, -1).view(np.uint8)[:, :n])
    # ETEXT: word k picks DICT[mix64(sb + (k+1)*GOLD) >> 61]; laid out sequentially, truncated at n
    out = np.empty((n_blocks, n), dtype=np.uint8)
    lens = np.array([len(w) + 1 for w in DICT], dtype=np.int64)
    flat = np.frombuffer(b"".join(w + b" " for w in DICT), dtype=np.uint8)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    n_words = n // 3 + 2                                        # shortest entry is 3 bytes
    pick = (_words(sb, n_words) >> np.uint64(61)).astype(np.int64)
    for b in range(n_blocks):
        l = lens[pick[b]]
        end = np.cumsum(l)
        cnt = int(np.searchsorted(end, n, side="left")) + 1
        l, p, e = l[:cnt], pick[b, :cnt], end[:cnt]
        src_idx = np.repeat(starts[p] - (e - l), l) + np.arange(int(e[-1]))
        out[b] = flat[src_idx[:n]]
    return out


def fuz_block(seed: int, length: int = 32768) -> np.ndarray:
    """The generator of the upstream fuzzer (original/fuzzer.c:81-85,153-168), one buffer per seed.

    Small sizes only (pure-Python loop).  Used by the size +-1 invariants tests.
    """
    P1, P2, P3, MASK = 2654435761, 2246822519, 3266489917, 0xFFFFFFFF
    st = [seed & MASK]

    def rnd():
        st[0] = (st[0] * P1 + P2) & MASK
        return st[0]

    rnd()
    seeds = []
    for _ in range(4):
        v = (rnd() << 8) & MASK
        v ^= (rnd() >> 8) & 65535
        seeds.append(v)
    cur = P3
    out = np.empty(length, dtype=np.uint8)
    for j in range(length):
        k = rnd()
        if j == 0 or ((k >> 10) % 10) == 0:
            cur = seeds[(rnd() >> 16) & 3]
        if ((k >> 8) & 255) == 0:
            i = (rnd() >> 16) & 3
            v = (rnd() << 8) & MASK
            v ^= (rnd() >> 8) & 65535
            seeds[i] = v
        cur = (cur * P1 + P2) & MASK
        out[j] = (cur >> 16) & 255
    return out
