"""HC encoder quick check: device-timed HC encode of one batch per class (GB/s raw) + round trip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lz4net_b200
from bench import Workload, BLOCK, GB
ctx = lz4net_b200.Context(0)
for cls, nb in (("E50", 131072), ("ETEXT", 32768)):
    w = Workload(ctx, nb, cls, nb, seed=3)
    w.encode(hc=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w.encode(hc=True); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    for wv in range(w.n_waves):
        b0, b1 = w.decode_wave(wv); torch.cuda.synchronize()
        assert torch.equal(w.out[: (b1 - b0) * BLOCK], w.raw[b0 * BLOCK: b1 * BLOCK])
    print(cls, "HC GB/s", round(nb * BLOCK / t / GB, 2), "ratio", round(int(w.clen.sum()) / (nb * BLOCK), 4), flush=True)
    del w; torch.cuda.empty_cache()
