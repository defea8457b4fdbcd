"""A/B of encode_stream_sharded's piece count with a host-side timeline of every rank (run under torch.distributed.run).
Usage: python -m torch.distributed.run --nproc-per-node 2 tools/stream_ab.py [GiB]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import lz4net_b200
from lz4net_b200 import batch, shard, synth

BLOCK = 65536
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8
nb = int(gib * (1 << 30)) // BLOCK
ctx = lz4net_b200.Context(lr)
enc0, dec = shard.gpu_codec(ctx, BLOCK)
raw = None
if rank == 0:
    raw = torch.empty(nb * BLOCK, dtype=torch.uint8, device=dev)
    for b0 in range(0, nb, 65536):
        batch.synth_fill(ctx, raw[b0 * BLOCK:], min(65536, nb - b0), BLOCK, synth.CLASS_ID["E50"], seed=6, first_block=b0)
log = []
t0 = [0.0]
def enc(buf, m):
    return enc0(buf, m)
def report(tag, pieces, t, td, ok=None):
    ts = [None] * world
    dist.all_gather_object(ts, (t, list(log), td))
    if rank == 0:
        tm = max(x[0] for x in ts); tdm = max(x[2] for x in ts)
        print(json.dumps({"mode": tag, "pieces": pieces, "ms": round(tm * 1e3, 1), "gbs": round(nb * BLOCK / tm / 1e9, 1), "dec_ms": round(tdm * 1e3, 1),
                          "dec_gbs": round(nb * BLOCK / tdm / 1e9, 1), "ok": ok, "timeline_ms(start,end,blocks)": [x[1] for x in ts]}), flush=True)

for pieces in (1,):
    for it in range(2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        log.clear(); t0[0] = time.perf_counter()
        lens, off, packed = shard.encode_stream_sharded(raw, nb, BLOCK, enc, rank, world, device=dev, pieces=pieces)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0[0]
        dist.barrier(); torch.cuda.synchronize()
        d0 = time.perf_counter()
        back = shard.decode_stream_sharded(packed, lens, nb, BLOCK, dec, rank, world, device=dev)
        torch.cuda.synchronize()
        td = time.perf_counter() - d0
        del back
    report("send/recv", pieces, t, td)
want = packed.clone() if rank == 0 else None
del packed, lens, off
torch.cuda.empty_cache()
w0 = time.perf_counter()
win = shard.StreamWindow(nb, BLOCK, rank, world, device=dev)
if rank == 0:
    win.raw.copy_(raw); print("window set-up ms", round((time.perf_counter() - w0) * 1e3, 1), flush=True)
for impl, pieces in (("capi", 1), ("capi", 2), ("capi", 4), ("capi", 8)):
    for it in range(2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        log.clear(); t0[0] = time.perf_counter()
        tr = []
        lens, off, packed = shard.encode_stream_window(win, enc, pieces=pieces, trace=tr)
        t = time.perf_counter() - t0[0]
        torch.cuda.synchronize()
        log[:] = [(k, round(tr[0][1].elapsed_time(e), 1)) for k, e in tr[1:]]
        ok = None
        if rank == 0:
            ok = bool(torch.equal(packed, want)); win.raw.zero_()
        dist.barrier(); torch.cuda.synchronize()
        d0 = time.perf_counter()
        tr = []
        back = shard.decode_stream_window(win, dec, pieces=pieces, trace=tr)
        td = time.perf_counter() - d0
        torch.cuda.synchronize()
        log += [("|", 0)] + [(k, round(tr[0][1].elapsed_time(e), 1)) for k, e in tr[1:]]
        if rank == 0:
            ok = ok and bool(torch.equal(back, raw))
    report("window/" + impl, pieces, t, td, ok)
win.close()
dist.barrier()
dist.destroy_process_group()
