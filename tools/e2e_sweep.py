"""End-to-end (host buffers through the C ABI) sweep over the pipeline chunk size, next to the box's raw PCIe rates.
usage: python tools/e2e_sweep.py [class] [blocks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lz4net_b200
from bench import e2e_host, GB
cls = sys.argv[1] if len(sys.argv) > 1 else "E50"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 14
# raw PCIe: 1 GiB pinned <-> device, both directions, alone and together
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h2 = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    th = timed(lambda: d.copy_(h, non_blocking=True)); td = timed(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
tb = timed(both)
print(f"pcie h2d {(1 << 30) / th / GB:.1f} GB/s  d2h {(1 << 30) / td / GB:.1f} GB/s  both at once {(1 << 30) / tb / GB:.1f} GB/s each", flush=True)
del h, d, h2, d2
ctx = lz4net_b200.Context(0)
for mb in (32, 64, 128, 256, 512):
    ctx.set_option("host_chunk_mb", mb)
    r = e2e_host(ctx, cls, nb, 3, 1)
    print("chunk_mb", mb, "enc", round(r["bytes"] / r["t_enc"] / GB, 1), "dec", round(r["bytes"] / r["t_dec"] / GB, 1),
          "rt", round(r["bytes"] / (r["t_enc"] + r["t_dec"]) / GB, 1), flush=True)
