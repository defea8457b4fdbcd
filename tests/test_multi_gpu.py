"""Two GPUs, NCCL: BASELINE configs[3] in miniature -- one raw stream on rank 0, blocks scattered over the ranks, encoded,
payloads gathered back in stream order; then the mirror image.  The result must be byte-identical to encoding the whole
stream on one GPU (blocks are independent: doc/compatibility.md:4-7).  Skipped on a box with fewer than two GPUs."""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_blocks, out_dir):
    import torch
    import torch.distributed as dist
    import lz4net_b200
    from lz4net_b200 import batch, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = lz4net_b200.Context(rank)
    bs = 65536
    enc, dec = shard.gpu_codec(ctx, bs)
    raw = None
    if rank == 0:
        raw = torch.empty(n_blocks * bs, dtype=torch.uint8, device="cuda")
        half = n_blocks // 2
        batch.synth_fill(ctx, raw, half, bs, synth.CLASS_ID["E50"], seed=4)
        batch.synth_fill(ctx, raw[half * bs:], n_blocks - half, bs, synth.CLASS_ID["ETEXT"], seed=4, first_block=half)
        torch.cuda.synchronize()
    lens, off, packed = shard.encode_stream_sharded(raw, n_blocks, bs, enc, rank, world, device=torch.device("cuda", rank))
    if rank == 0:
        want_p, want_l = enc(raw, n_blocks)                        # the whole stream on one GPU
        assert torch.equal(lens, want_l) and torch.equal(packed, want_p)
    back = shard.decode_stream_sharded(packed, lens, n_blocks, bs, dec, rank, world, device=torch.device("cuda", rank))
    if rank == 0:
        assert torch.equal(back, raw)
    # the same over peer memory (CUDA IPC window on the root, copy-engine transfers): same bytes
    win = shard.StreamWindow(n_blocks, bs, rank, world, device=torch.device("cuda", rank))
    if rank == 0:
        win.raw.copy_(raw)
    wl, wo, wp = shard.encode_stream_window(win, enc, pieces=3)
    if rank == 0:
        assert torch.equal(wl, lens) and torch.equal(wp, packed) and int(wo[-1]) == packed.numel()
        win.raw.zero_()
    wb = shard.decode_stream_window(win, dec, pieces=3)
    if rank == 0:
        assert torch.equal(wb, raw)
        open(os.path.join(out_dir, "ok"), "w").write("1")
    win.close()
    dist.barrier()
    dist.destroy_process_group()


def test_stream_sharded_over_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), 1031, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
