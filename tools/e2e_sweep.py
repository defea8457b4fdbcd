import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lz4net_b200
from bench import e2e_host, GB
ctx = lz4net_b200.Context(0)
for mb in (96, 256, 512, 1024):
    ctx.set_option("host_chunk_mb", mb)
    r = e2e_host(ctx, "E50", 1 << 15, 3, 1)
    print("chunk_mb", mb, "enc", round(r["bytes"] / r["t_enc"] / GB, 1), "dec", round(r["bytes"] / r["t_dec"] / GB, 1),
          "rt", round(r["bytes"] / (r["t_enc"] + r["t_dec"]) / GB, 1), flush=True)
