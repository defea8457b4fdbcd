"""Host-side logic of the LZ4Stream mirror (lz4net_b200/codec.py) without a GPU: the class only asks its context for
stream_encode / stream_decode of whole buffers, so a stand-in context built on the ORACLE exercises everything else --
chunk boundaries, InteractiveRead (src/LZ4/LZ4Stream.cs:376-401), the byte caps on read-ahead and write buffer, the
length checks of AcquireNextChunk (:274-312).  The GPU twin is tests/test_gpu_parity.py."""
import io

import pytest

import oracle
from lz4net_b200 import codec
from tests import cases


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F; v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _read_varint(buf, p):
    v = sh = 0
    while True:
        b = buf[p]; p += 1; v |= (b & 0x7F) << sh; sh += 7
        if not b & 0x80:
            return v, p


class OracleContext:
    """stream_encode / stream_decode as the reference's LZ4Stream frames them (FlushCurrentChunk :239-269)."""

    def stream_encode(self, data, bs, hc):
        out = bytearray()
        for o in range(0, len(data), bs):
            blk = data[o:o + bs]
            r, c = (oracle.encode_hc if hc else oracle.encode)(blk, cap=len(blk))
            if r <= 0 or r >= len(blk):
                out += _varint(0) + _varint(len(blk)) + blk
            else:
                out += _varint(1) + _varint(len(blk)) + _varint(r) + c[:r]
        return bytes(out)

    def stream_decode(self, raw):
        out, p = bytearray(), 0
        while p < len(raw):
            f, p = _read_varint(raw, p); rl, p = _read_varint(raw, p)
            cl, p = _read_varint(raw, p) if f & 1 else (rl, p)
            if f & 1:
                r, o = oracle.decode_known(raw[p:p + cl], rl)
                assert r == cl
                out += o
            else:
                out += raw[p:p + cl]
            p += cl
        return bytes(out)


def _chunk_ends(wire):
    ends, pos = [], 0
    while pos < len(wire):
        f, pos = _read_varint(wire, pos); rl, pos = _read_varint(wire, pos)
        cl, pos = _read_varint(wire, pos) if f & 1 else (rl, pos)
        pos += cl; ends.append(pos)
    return ends


S, M, F = codec.LZ4Stream, codec.LZ4StreamMode, codec.LZ4StreamFlags
BS = 4096


@pytest.fixture(scope="module")
def stream():
    ctx = OracleContext()
    data = (cases.content("ETEXT", 7 * BS, seed=9).tobytes() + cases.content("E0", 2 * BS, seed=9).tobytes()
            + cases.content("E100", BS + 100, seed=9).tobytes())
    wire = ctx.stream_encode(data, BS, False)
    return ctx, data, wire, _chunk_ends(wire)


def test_interactive_read_takes_one_chunk_at_a_time(stream):
    ctx, data, wire, ends = stream
    inner = io.BytesIO(wire)
    r = S(inner, M.Decompress, F.InteractiveRead, batchBlocks=256, context=ctx)
    back = bytearray()
    for k in range(len(ends)):
        got = r.Read(1 << 20)
        assert inner.tell() == ends[k] and len(got) == min(BS, len(data) - k * BS)    # one chunk read, one chunk returned
        back += got
    assert r.Read(10) == b"" and bytes(back) == data
    r = S(io.BytesIO(wire), M.Decompress, F.InteractiveRead, context=ctx)              # short reads inside a chunk
    assert r.Read(100) == data[:100] and r.Read(BS) == data[100:BS] and r.ReadByte() == data[BS]


def test_batched_read_is_capped_by_chunks_and_by_bytes(stream):
    ctx, data, wire, ends = stream
    inner = io.BytesIO(wire)
    r = S(inner, M.Decompress, batchBlocks=4, context=ctx)
    assert len(r.Read(1)) == 1 and inner.tell() == ends[3]
    inner = io.BytesIO(wire)
    r = S(inner, M.Decompress, batchBlocks=256, context=ctx, maxBufferBytes=3 * BS)
    assert len(r.Read(1)) == 1 and inner.tell() == ends[2]
    assert r.Read(len(data)) == data[1:] and r.Read(1) == b""
    inner = io.BytesIO(wire)
    r = S(inner, M.Decompress, batchBlocks=256, context=ctx, maxBufferBytes=1)            # never less than one chunk
    assert r.Read(BS + 1) == data[:BS + 1] and inner.tell() == ends[1]


@pytest.mark.parametrize("hc", [False, True])
def test_write_buffer_cap_and_chunk_boundaries(stream, hc):
    ctx, data, _, _ = stream
    wire = ctx.stream_encode(data, BS, hc)
    ends = _chunk_ends(wire)
    out = io.BytesIO()
    flags = F.IsolateInnerStream | (F.HighCompression if hc else 0)
    w = S(out, M.Compress, flags, BS, batchBlocks=256, context=ctx, maxBufferBytes=2 * BS)
    w.Write(data[:2 * BS]); assert out.tell() == 0                  # a full buffer waits for more data (:463-467)
    w.Write(data[2 * BS:2 * BS + 1]); assert out.tell() == ends[1]
    w.Write(data[2 * BS + 1:]); w.Close()
    assert out.getvalue() == wire and not out.closed               # IsolateInnerStream
    out = io.BytesIO()
    with S(out, M.Compress, F.IsolateInnerStream, BS, batchBlocks=3, context=ctx) as w:  # Flush ends the current chunk
        w.Write(data[:BS + 10]); w.Flush(); w.Write(data[BS + 10:])
    assert out.getvalue() == ctx.stream_encode(data[:BS + 10], BS, False) + ctx.stream_encode(data[BS + 10:], BS, False)


def test_truncated_and_corrupt_streams(stream):
    ctx, data, wire, ends = stream
    with pytest.raises(EOFError):
        S(io.BytesIO(wire[:-2]), M.Decompress, context=ctx).Read(len(data) + 1)
    with pytest.raises(EOFError):
        S(io.BytesIO(wire[:ends[0] + 1]), M.Decompress, context=ctx).Read(len(data) + 1)   # a header cut after its flags
    bad = _varint(1) + _varint(10) + _varint(11) + bytes(11)                               # compressed longer than raw (:288)
    with pytest.raises(EOFError):
        S(io.BytesIO(bad), M.Decompress, context=ctx).Read(1)
    huge = _varint(0) + _varint(1 << 40)                                                   # a length no int holds
    with pytest.raises(EOFError):
        S(io.BytesIO(huge), M.Decompress, context=ctx).Read(1)
    assert S(io.BytesIO(b""), M.Decompress, context=ctx).Read(5) == b""
    with pytest.raises(NotImplementedError):
        S(io.BytesIO(), M.Compress, context=ctx).Read(1)
    with pytest.raises(NotImplementedError):
        S(io.BytesIO(wire), M.Decompress, context=ctx).Write(b"x")
