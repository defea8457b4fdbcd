"""CPU-only, world_size = 2 over gloo: the N > 1 plumbing of bench.py / lz4net_b200.shard (block ownership, the
max-over-ranks timing rule, whole-job aggregation) without a GPU.  Rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lz4net_b200 import shard, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, blocks_per_rank, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.weak_range(rank, blocks_per_rank)
    data = synth.make_blocks("E50", hi - lo, 4096, seed=9, first_block=lo)       # this rank's shard, by GLOBAL index
    np.save(os.path.join(out_dir, f"shard{rank}.npy"), data)
    # pretend rank r needed (r+1) seconds for its shard: the job's time is the max, its bytes the sum
    secs = float(rank + 1)
    tput = shard.aggregate_throughput(float(data.size), secs)
    worst, = shard.reduce_max([secs])
    total, = shard.reduce_sum([float(data.size)])
    assert worst == float(world) and total == float(world * data.size)
    assert abs(tput - total / worst) < 1e-9
    # strong split: ranges are contiguous, ordered, cover [0, N)
    owned = torch.zeros(37, dtype=torch.int32)
    a, b = shard.strong_range(rank, world, 37)
    owned[a:b] += 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world, bpr = 2, 6
    mp.spawn(_worker, args=(world, _free_port(), bpr, str(tmp_path)), nprocs=world, join=True)
    whole = synth.make_blocks("E50", world * bpr, 4096, seed=9, first_block=0)
    got = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    assert np.array_equal(got, whole)            # shards are exactly the global batch, in stream order


def test_ranges():
    assert shard.weak_range(3, 10) == (30, 40)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64, 1000):
            rs = [shard.strong_range(r, world, n) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


# ---- one stream, many ranks (BASELINE configs[3]): scatter from the root, encode per rank, gather in stream order ------
def _toy_codec(bs):
    """A stand-in codec for the CPU test of the plumbing: a block is 'compressed' by dropping its trailing zero bytes."""
    def enc(raw, m):
        blocks = raw.view(m, bs)
        lens = torch.tensor([max(1, int(torch.nonzero(b).max().item()) + 1 if bool(b.any()) else 1) for b in blocks], dtype=torch.int32)
        return torch.cat([b[:int(n)] for b, n in zip(blocks, lens)]), lens

    def dec(packed, lens, m, out=None):
        out = torch.zeros(m * bs, dtype=torch.uint8) if out is None else out.zero_()
        p = 0
        for i, n in enumerate(lens.tolist()):
            out[i * bs:i * bs + n] = packed[p:p + n]; p += n
        return out
    return enc, dec


def _stream_worker(rank, world, port, n_blocks, bs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    enc, dec = _toy_codec(bs)
    raw = None
    if rank == 0:
        g = torch.Generator().manual_seed(5)
        raw = torch.randint(0, 256, (n_blocks * bs,), dtype=torch.uint8, generator=g)
        for i in range(n_blocks):                                  # zero tails of different lengths -> different payload sizes
            raw[i * bs + (37 * i) % bs + 1:(i + 1) * bs] = 0
    lens, off, packed = shard.encode_stream_sharded(raw, n_blocks, bs, enc, rank, world)
    if rank == 0:
        want_p, want_l = enc(raw, n_blocks)                        # the same stream through one rank
        assert torch.equal(lens, want_l) and torch.equal(packed, want_p)
        assert off.tolist() == [0] + torch.cumsum(want_l.to(torch.int64), 0).tolist()
    back = shard.decode_stream_sharded(packed, lens, n_blocks, bs, dec, rank, world)
    if rank == 0:
        assert torch.equal(back, raw)
        open(os.path.join(out_dir, "ok"), "w").write("1")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_blocks", [(2, 7), (3, 2), (2, 1)])
def test_stream_scatter_encode_gather_gloo(tmp_path, world, n_blocks):
    """More ranks than blocks, uneven splits, one block: order and bytes are those of the single-rank result."""
    mp.spawn(_stream_worker, args=(world, _free_port(), n_blocks, 64, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def _window_refused_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard.StreamWindow(4, 64, rank, world)                     # no CUDA here: the root refuses, every rank hears about it
    except RuntimeError as e:
        assert "StreamWindow" in str(e)
        open(os.path.join(out_dir, f"refused{rank}"), "w").write("1")
    dist.barrier()
    dist.destroy_process_group()


def test_stream_window_failure_is_collective(tmp_path):
    """Peer memory needs GPUs; without them the set-up fails on EVERY rank with the root's reason instead of hanging."""
    mp.spawn(_window_refused_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "refused0").exists() and (tmp_path / "refused1").exists()
