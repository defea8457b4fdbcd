"""Multi-GPU sharding of a batch of independent blocks.

lz4net's blocks are independent by construction (doc/compatibility.md:4-7), so the multi-GPU form of the path is plain
data parallelism: rank r owns the contiguous block range [r*B, (r+1)*B) (weak scaling: B blocks per GPU) or an even
split of a fixed total (strong scaling, used by LZ4Stream-style callers to keep stream order on gather).  The codec has
no data-path collective; bench.py only uses torch.distributed for the barrier and the max-over-ranks of the timings.
When the data starts on one rank (BASELINE configs[3]: one stream, many GPUs) the second half of this module scatters
the blocks from the root and gathers the payloads back, in stream order.
"""
from __future__ import annotations

from typing import Sequence, Tuple


def weak_range(rank: int, blocks_per_rank: int) -> Tuple[int, int]:
    """Global block indices owned by `rank` when every rank processes `blocks_per_rank` blocks."""
    return rank * blocks_per_rank, (rank + 1) * blocks_per_rank


def strong_range(rank: int, world: int, total_blocks: int) -> Tuple[int, int]:
    """Contiguous, order-preserving split of `total_blocks` over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(total_blocks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(values: Sequence[float], device=None):
    """Max over ranks of each value (device timings: the slowest rank defines the job's time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def reduce_sum(values: Sequence[float], device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(bytes_per_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job bytes/s: all ranks' bytes over the slowest rank's time."""
    total = reduce_sum([bytes_per_rank], device)[0]
    worst = reduce_max([seconds_this_rank], device)[0]
    return total / worst


# ----------------------------------------------------------------------------------------------------------------------
# One stream, many GPUs (BASELINE configs[3]): the root rank holds a raw stream chunked into independent blocks; every
# rank encodes a contiguous, order-preserving range of them; the root ends up with every block's compressed length
# and the packed payloads in stream order (exactly what LZ4Stream's writer consumes, src/LZ4/LZ4Stream.cs:239-269).
# The decode direction is the mirror image.  The only exchange steps are the scatter of the inputs and the gather of the
# outputs (point-to-point sends between the root and each peer: NCCL over NVLink on a GPU box, gloo in the CPU tests);
# the codec itself never communicates.  `encode_local` / `decode_local` are the per-rank codec calls -- the GPU batch
# functions below in production, a toy codec in the CPU tests of the plumbing.
# ----------------------------------------------------------------------------------------------------------------------
def _p2p(ops):
    import torch.distributed as dist
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_ranges(root_buf, byte_ranges, rank: int, root: int = 0, like=None):
    """The root sends bytes [lo, hi) of `root_buf` to rank r for byte_ranges[r] = (lo, hi); returns this rank's bytes
    (a view on the root, a new tensor elsewhere).  Every rank passes the same byte_ranges."""
    import torch
    import torch.distributed as dist
    lo, hi = byte_ranges[rank]
    if rank == root:
        _p2p([dist.P2POp(dist.isend, root_buf[a:b], r) for r, (a, b) in enumerate(byte_ranges) if r != root and b > a])
        return root_buf[lo:hi]
    mine = torch.empty(hi - lo, dtype=torch.uint8, device=like.device if like is not None else "cpu")
    _p2p([dist.P2POp(dist.irecv, mine, root)] if hi > lo else [])
    return mine


def gather_ranges(local, byte_ranges, rank: int, root: int = 0, out=None):
    """Mirror of scatter_ranges: rank r's `local` bytes land at byte_ranges[r] of `out` on the root (returned there)."""
    import torch.distributed as dist
    if rank == root:
        lo, hi = byte_ranges[root]
        out[lo:hi] = local[: hi - lo]
        _p2p([dist.P2POp(dist.irecv, out[a:b], r) for r, (a, b) in enumerate(byte_ranges) if r != root and b > a])
        return out
    lo, hi = byte_ranges[rank]
    _p2p([dist.P2POp(dist.isend, local[: hi - lo], root)] if hi > lo else [])
    return None


def _all_gather_i64(value: int, device):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    return [int(t.item()) for t in out]


def encode_stream_sharded(raw_root, n_blocks: int, block_size: int, encode_local, rank: int, world: int, root: int = 0,
                          device=None, pieces: int = 4):
    """raw_root: uint8[n_blocks * block_size] on the root (ignored elsewhere; the last block may not be short).
    encode_local(shard_bytes, m_blocks) -> (packed uint8[...], lens int32[m_blocks]) for m_blocks blocks.
    Returns on the root (lens int32[n_blocks], offsets int64[n_blocks + 1], packed uint8[total]) in stream order,
    (None, None, None) elsewhere.

    Every rank's block range is cut into `pieces` sub-ranges; all transfers are posted up front and a piece is encoded as
    soon as it has arrived, so the scatter of piece p+1 (NCCL's stream) overlaps the encode of piece p (the compute
    stream): the root's NVLink egress and the kernels work at the same time instead of one after the other."""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else (raw_root.device if raw_root is not None else "cpu")
    blk = [strong_range(r, world, n_blocks) for r in range(world)]
    like = torch.empty(0, dtype=torch.uint8, device=dev)
    a, b = blk[rank]
    m = b - a
    pieces = max(1, min(pieces, m)) if m else 1
    def sub(r, p, k):                                           # block sub-range p of k of rank r
        lo, hi = blk[r]; n = hi - lo
        return lo + n * p // k, lo + n * (p + 1) // k
    works, bufs = [], []
    for p in range(pieces):
        if rank == root:
            ops = []
            for r in range(world):
                if r == root:
                    continue
                k = max(1, min(pieces, blk[r][1] - blk[r][0])) if blk[r][1] > blk[r][0] else 1
                if p < k:
                    lo, hi = sub(r, p, k)
                    if hi > lo:
                        ops.append(dist.P2POp(dist.isend, raw_root[lo * block_size: hi * block_size], r))
            works.append(dist.batch_isend_irecv(ops) if ops else [])
            lo, hi = sub(root, p, pieces) if m else (0, 0)
            bufs.append(raw_root[lo * block_size: hi * block_size])
        else:
            lo, hi = sub(rank, p, pieces) if m else (0, 0)
            buf = torch.empty((hi - lo) * block_size, dtype=torch.uint8, device=dev)
            works.append(dist.batch_isend_irecv([dist.P2POp(dist.irecv, buf, root)]) if hi > lo else [])
            bufs.append(buf)
    packed_parts, lens_parts = [], []
    for p in range(pieces):
        if rank != root:
            for w in works[p]:
                w.wait()
        lo, hi = sub(rank, p, pieces) if m else (0, 0)
        if hi > lo:
            pk, ln = encode_local(bufs[p], hi - lo)
            packed_parts.append(pk); lens_parts.append(ln)
    if rank == root:
        for ws in works:
            for w in ws:
                w.wait()
    packed = torch.cat(packed_parts) if len(packed_parts) > 1 else (packed_parts[0] if packed_parts else like)
    lens = torch.cat(lens_parts) if len(lens_parts) > 1 else (lens_parts[0] if lens_parts else torch.empty(0, dtype=torch.int32, device=dev))
    total = int(lens.to(torch.int64).sum().item()) if m else 0
    totals = _all_gather_i64(total, dev)                        # every rank's payload size: the only metadata exchanged
    pay = [(sum(totals[:r]), sum(totals[:r + 1])) for r in range(world)]
    out = torch.empty(sum(totals), dtype=torch.uint8, device=dev) if rank == root else None
    out = gather_ranges(packed, pay, rank, root, out)
    lens_all = torch.empty(n_blocks * 4, dtype=torch.uint8, device=dev) if rank == root else None
    lens_all = gather_ranges(lens.contiguous().view(torch.uint8) if m else like, [(x * 4, y * 4) for x, y in blk], rank, root, lens_all)
    if rank != root:
        return None, None, None
    lens_i = lens_all.view(torch.int32)
    off = torch.zeros(n_blocks + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens_i.to(torch.int64), 0)
    return lens_i, off, out


def decode_stream_sharded(packed_root, lens_root, n_blocks: int, block_size: int, decode_local, rank: int, world: int,
                          root: int = 0, device=None):
    """Mirror image: packed_root / lens_root (int32[n_blocks]) on the root; decode_local(packed, lens, m_blocks) ->
    uint8[m_blocks * block_size].  Returns the raw stream on the root, None elsewhere."""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else (packed_root.device if packed_root is not None else "cpu")
    blk = [strong_range(r, world, n_blocks) for r in range(world)]
    like = torch.empty(0, dtype=torch.uint8, device=dev)
    lens = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    if rank == root:
        lens.copy_(lens_root)
    dist.broadcast(lens, src=root)                               # 4 bytes per block: everybody derives the payload ranges
    off = torch.zeros(n_blocks + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens.to(torch.int64), 0)
    offs = off.tolist()
    pay = [(offs[a], offs[b]) for a, b in blk]
    mine = scatter_ranges(packed_root, pay, rank, root, like=like)
    a, b = blk[rank]
    raw = decode_local(mine, lens[a:b].contiguous(), b - a) if b > a else like
    out = torch.empty(n_blocks * block_size, dtype=torch.uint8, device=dev) if rank == root else None
    return gather_ranges(raw, [(x * block_size, y * block_size) for x, y in blk], rank, root, out)


def gpu_codec(ctx, block_size: int, hc: bool = False):
    """(encode_local, decode_local) on device tensors through the C ABI (lz4b200_encode_batch / compact / decode_batch)."""
    import torch
    from . import batch

    def encode_local(raw, m):
        slot = block_size + block_size // 255 + 16
        so, do, sl, dc = batch.uniform_layout(m, block_size, slot, raw.device)
        slots = torch.empty(m * slot, dtype=torch.uint8, device=raw.device)
        lens = torch.zeros(m, dtype=torch.int32, device=raw.device)
        batch.encode(ctx, raw, so, sl, slots, do, dc, lens, hc=hc)
        off = torch.zeros(m + 1, dtype=torch.int64, device=raw.device)
        packed = torch.empty(m * slot, dtype=torch.uint8, device=raw.device)
        batch.compact(ctx, slots, do, lens, packed, off)
        torch.cuda.current_stream().synchronize()               # (this stream only: transfers of later pieces keep running)
        assert int((lens <= 0).sum()) == 0, "encode failed"
        return packed[: int(off[-1].item())], lens

    def decode_local(packed, lens, m):
        off = torch.zeros(m + 1, dtype=torch.int64, device=packed.device)
        off[1:] = torch.cumsum(lens.to(torch.int64), 0)
        so, _, sl, _ = batch.uniform_layout(m, block_size, block_size, packed.device)
        out = torch.empty(m * block_size, dtype=torch.uint8, device=packed.device)
        used = torch.zeros(m, dtype=torch.int32, device=packed.device)
        batch.decode(ctx, packed, off[:-1].contiguous(), lens, out, so, sl, used, known=True)
        torch.cuda.current_stream().synchronize()
        assert torch.equal(used, lens), "decode rejected a stream"
        return out

    return encode_local, decode_local
