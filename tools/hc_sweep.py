"""HC encoder throughput vs blocks in flight (one thread per block)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lz4net_b200
from bench import Workload, BLOCK, GB
ctx = lz4net_b200.Context(0)
for cls, nb in (("E50", 131072), ("ETEXT", 65536)):
    w = Workload(ctx, nb, cls, nb, seed=3)
    for conc in (16384, 65536, 131072, 262144):
        if conc > nb * 2: continue
        ctx.set_option("hc_concurrency", conc)
        w.encode(hc=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); w.encode(hc=True); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        print(cls, "conc", conc, "GB/s", round(nb * BLOCK / t / GB, 2), "ratio", round(int(w.clen.sum()) / (nb * BLOCK), 4), flush=True)
    del w; torch.cuda.empty_cache()
