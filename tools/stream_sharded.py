"""BASELINE configs[3]: one synthetic stream on rank 0, chunked into 64 KiB blocks, encoded on all ranks, payloads gathered
back on rank 0 in stream order (and the mirror image).  Launch with torchrun; prints one JSON line on rank 0.
usage: torchrun --nproc-per-node N tools/stream_sharded.py [GiB=16] [class=E50]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lz4net_b200
from lz4net_b200 import batch, shard, synth

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
cls = sys.argv[2] if len(sys.argv) > 2 else "E50"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = lz4net_b200.Context(local)
bs = 65536
nb = int(gib * (1 << 30)) // bs
enc, dec = shard.gpu_codec(ctx, bs)
raw = None
if rank == 0:
    raw = torch.empty(nb * bs, dtype=torch.uint8, device=dev)
    for b0 in range(0, nb, 65536):
        batch.synth_fill(ctx, raw[b0 * bs:], min(65536, nb - b0), bs, synth.CLASS_ID[cls], seed=6, first_block=b0)
res = {}
for it in range(2):                                                 # first pass warms NCCL's connections up
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    lens, off, packed = shard.encode_stream_sharded(raw, nb, bs, enc, rank, world, device=dev)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    back = shard.decode_stream_sharded(packed, lens, nb, bs, dec, rank, world, device=dev)
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    res = {"encode_s": t1 - t0, "decode_s": t2 - t1}
if rank == 0:
    assert torch.equal(back, raw)
    n = nb * bs
    print(json.dumps({"workload": f"{gib:g} GiB stream of class {cls}, 64 KiB blocks, root scatter -> encode on {world} GPU(s) -> gather (and back)",
                      "n_gpus": world, "ratio": round(int(packed.numel()) / n, 4),
                      "encode_gbs": round(n / res["encode_s"] / 1e9, 2), "decode_gbs": round(n / res["decode_s"] / 1e9, 2),
                      "roundtrip_gbs": round(n / (res["encode_s"] + res["decode_s"]) / 1e9, 2)}), flush=True)
dist.destroy_process_group()
