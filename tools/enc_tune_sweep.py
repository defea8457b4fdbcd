"""Device-timed sweep of the encoder's path-selection heuristics (they never change the emitted bytes).
usage: python tools/enc_tune_sweep.py [blocks] [classes]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lz4net_b200
from bench import Workload, BLOCK, GB

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
CLASSES = sys.argv[2].split(",") if len(sys.argv) > 2 else ("E50", "ETEXT")
ctx = lz4net_b200.Context(0)
GRID = [("default", 12, 8, 24), ("lane32", 32, 8, 24), ("lane64", 64, 8, 24), ("probe0", 12, 0, 24), ("probe16", 12, 16, 24), ("probe4", 12, 4, 24),
        ("wide0", 12, 8, 0), ("wide16", 12, 8, 16), ("wide1000", 12, 8, 1000), ("lane32probe16", 32, 16, 24)]
for cls in CLASSES:
    w = Workload(ctx, nb, cls, nb, seed=2)
    rb = nb * BLOCK
    ref = None
    for name, a, b, c in GRID:
        ctx.set_option("encode_lane_copy_max", a); ctx.set_option("encode_probe_max", b); ctx.set_option("encode_wide_min", c)
        w.slots.zero_(); w.encode(); torch.cuda.synchronize()
        if ref is None:
            w.verify(); ref = (w.clen.clone(), w.slots.clone())
        else:
            assert torch.equal(w.clen, ref[0]) and torch.equal(w.slots, ref[1]), (cls, name, "bytes changed")
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); w.encode(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
        print(json.dumps({"cls": cls, "tune": name, "enc_gbs": round(rb / sorted(ts)[1] / GB, 1)}), flush=True)
    del w, ref; torch.cuda.empty_cache()
