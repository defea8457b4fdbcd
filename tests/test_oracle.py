"""CPU-only: pins the port oracle (oracle/lz4_oracle.c) to the reference's own sources (oracle/_ref) and to the
committed golden vectors.  Mirrors src/LZ4.Tests/ConformanceTests.cs (all encoders byte-identical, all decoders
round-trip) and original/fuzzer.c:146-233 (size +-1 invariants)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from tests import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def test_bound():
    for n in (0, 1, 254, 255, 256, 65536, 1 << 20):
        assert oracle.bound(n) == oracle.port().lz4o_bound(n) == n + n // 255 + 16
    assert oracle.bound(65536) == 65809            # src/LZ4/LZ4Codec.cs:313-316


def test_golden_vectors_port():
    """The port reproduces every committed golden vector (made from oracle/_ref by make_golden.py)."""
    g = json.load(open(GOLD))
    assert len(g["cases"]) >= 40
    for c in g["cases"]:
        data = cases.AUTOTEST if c["model"] == "autotest" else cases.content(c["model"], c["n"], c["seed"]).tobytes()
        assert _sha(data) == c["input_sha256"], c["name"]
        for mode, fn in (("fast", oracle.encode), ("hc", oracle.encode_hc)):
            r, out = fn(data, impl="port")
            assert r == c[mode]["len"], (c["name"], mode)
            assert _sha(out) == c[mode]["sha256"], (c["name"], mode)
            if "hex" in c[mode]:
                assert out.hex() == c[mode]["hex"]
            rr, dec = oracle.decode_known(out, len(data), impl="port")
            assert rr == r and dec == data
            # LZ4Stream / Wrap pass cap = n (src/LZ4/LZ4Stream.cs:243-246): the stored-raw decision is part of parity
            r2, _ = fn(data, cap=len(data), impl="port")
            assert r2 == c[mode]["len_cap_n"], (c["name"], mode)


@needs_ref
@pytest.mark.parametrize("model", cases.MODELS)
def test_port_equals_reference_random_lengths(model):
    lens = cases.random_lengths(60, 200_000, seed=7) + list(cases.BOUNDARY_LENGTHS)
    for i, n in enumerate(lens):
        if model in ("mixed",) and n > 70_000:
            n = n % 70_000
        data = cases.content(model, n, seed=i).tobytes()
        for fn in (oracle.encode, oracle.encode_hc):
            if fn is oracle.encode_hc and n > 70_000 and model in ("E100", "runs", "periodic"):
                pass
            rr, ro = fn(data, impl="ref")
            rp, po = fn(data, impl="port")
            assert (rr, ro) == (rp, po), (model, n, fn.__name__)
            # limited output: exactly enough, one short, and cap = n (fuzzer.c:212-227; LZ4Stream.cs:243-246)
            for cap in {rr, rr - 1, n, max(0, n - 1), rr // 2}:
                if cap < 0:
                    continue
                a = fn(data, cap=cap, impl="ref")
                b = fn(data, cap=cap, impl="port")
                assert a == b, (model, n, cap, fn.__name__)
            d1 = oracle.decode_known(ro, n, impl="port")
            d2 = oracle.decode_unknown(ro, n, impl="port")
            assert d1 == (rr, data) and d2 == (n if n else d2[0], data)


@needs_ref
def test_decoders_accept_reject_like_reference():
    """fuzzer.c:176-210: exact size works, size +-1 must fail; the port takes the same decisions as the reference."""
    rng = np.random.default_rng(3)
    from lz4net_b200 import synth
    for i in range(48):
        fuz = i % 6 == 5      # the upstream generator: for it the +-1 invariants are strict (fuzzer.c:176-210)
        data = (synth.fuz_block(i, 4096) if fuz else
                cases.content(cases.MODELS[i % len(cases.MODELS)], int(rng.integers(20, 40000)), seed=100 + i)).tobytes()
        n = len(data)
        _, comp = oracle.encode(data, impl="ref")
        clen = len(comp)
        for osize in (n, n - 1, n + 1):
            a = oracle.decode_known(comp, osize, impl="ref")[0]
            b = oracle.decode_known(comp, osize, impl="port")[0]
            assert (a < 0) == (b < 0) and (a < 0 or a == b), (i, osize, a, b)
            assert (osize == n) == (a >= 0)
        for isz, osz in ((clen, n + 1), (clen, n), (clen, n - 1), (clen - 1, n), (clen + 1, n)):
            cc = comp if isz <= clen else comp + b"\x00"
            a = oracle.decode_unknown(cc[:isz], osz, impl="ref")
            b = oracle.decode_unknown(cc[:isz], osz, impl="port")
            assert (a[0] < 0) == (b[0] < 0), (i, isz, osz, a[0], b[0])
            if a[0] >= 0:
                assert a == b
            if fuz or isz == clen:
                # (on arbitrary data a truncated stream can, rarely, still parse: the match-length loop stops at
                #  iend-6 and re-reads its last byte as a token, original/lz4.c:986-999)
                assert (a[0] >= 0) == (isz == clen and osz >= n)


@needs_ref
def test_corrupt_streams_same_verdict():
    """Bit-flipped streams: the port must never crash and must agree with the reference on accept/reject
    (and on the bytes when both accept), except for offset-0 matches which the port rejects by design."""
    rng = np.random.default_rng(11)
    agree = 0
    for i in range(300):
        data = cases.content("mixed", 3000, seed=i).tobytes()
        _, comp = oracle.encode(data, impl="ref")
        c = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            c[int(rng.integers(0, len(c)))] ^= 1 << int(rng.integers(0, 8))
        a = oracle.decode_unknown(bytes(c), len(data), impl="ref")
        b = oracle.decode_unknown(bytes(c), len(data), impl="port")
        if (a[0] < 0) == (b[0] < 0):
            agree += 1
            if a[0] >= 0:
                assert a == b
        else:
            assert a[0] >= 0 and b[0] < 0          # only the documented tightening (offset == 0)
    assert agree >= 290


def test_fuz_generator_roundtrip():
    from lz4net_b200 import synth
    for seed in range(4):
        data = synth.fuz_block(seed, 8192).tobytes()
        r, c = oracle.encode(data)
        rh, ch = oracle.encode_hc(data)
        assert oracle.decode_known(c, len(data)) == (r, data)
        assert oracle.decode_known(ch, len(data)) == (rh, data)
        assert oracle.decode_known(c, len(data) - 1)[0] < 0 and oracle.decode_known(c, len(data) + 1)[0] < 0
