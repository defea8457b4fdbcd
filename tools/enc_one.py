"""Encode one synthetic batch a few times (profiling target for ncu).
usage: python tools/enc_one.py [blocks] [class] [variant] [prefetch] [warps] [lane_warp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lz4net_b200
from bench import Workload, BLOCK

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cls = sys.argv[2] if len(sys.argv) > 2 else "E50"
ctx = lz4net_b200.Context(0)
if len(sys.argv) > 3: ctx.set_option("encode_variant", int(sys.argv[3]))
if len(sys.argv) > 4: ctx.set_option("encode_prefetch", int(sys.argv[4]))
if len(sys.argv) > 5: ctx.set_option("encode_ctas_per_sm", int(sys.argv[5]))
w = Workload(ctx, nb, cls, nb, seed=2)
for _ in range(3):
    w.encode()
torch.cuda.synchronize()
w.verify()
print("ok")
