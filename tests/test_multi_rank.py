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
