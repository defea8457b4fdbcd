"""Decode one synthetic batch a few times (profiling target for ncu).
usage: python tools/dec_one.py [blocks] [class] [lanes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lz4net_b200
from bench import Workload, BLOCK

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cls = sys.argv[2] if len(sys.argv) > 2 else "E50"
ctx = lz4net_b200.Context(0)
ctx.set_option("decode_lanes", int(sys.argv[3]) if len(sys.argv) > 3 else 108)
w = Workload(ctx, nb, cls, nb, seed=2)
w.encode()
for _ in range(3):
    w.decode_wave(0)
torch.cuda.synchronize()
assert torch.equal(w.out[: nb * BLOCK], w.raw[: nb * BLOCK])
print("ok")
