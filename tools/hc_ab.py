"""LZ4HC kernels side by side on one GPU: thread per block (hc_kernel 0) against the warp-per-block kernels on a static
index (1: block staged in shared memory, 2: block through L1) at several residencies.  Device-timed, one warm-up + one
timed pass per configuration; every configuration's lengths and a sample of its bytes are compared with kernel 0's
(the oracle parity is tests/test_gpu_parity.py's job).  Writes gpurun_out/hc_ab.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lz4net_b200
from bench import Workload, BLOCK, GB

T0 = time.time()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0          # seconds this tool may take
ctx = lz4net_b200.Context(0)
out = {"device": torch.cuda.get_device_name(0), "block": BLOCK, "runs": []}
CONFIGS = [(0, 0), (2, 16), (2, 32), (1, 3), (2, 24), (2, 8), (1, 2)]


def timed(w, kernel, warps):
    ctx.set_option("hc_kernel", kernel); ctx.set_option("hc_warps_per_sm", warps)
    w.clen.zero_()
    w.encode(hc=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w.encode(hc=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


for cls, nb in (("E50", 131072), ("ETEXT", 65536), ("E0", 32768)):
    w = Workload(ctx, nb, cls, nb, seed=3)
    ref_len = ref_bytes = None
    sample = list(range(0, nb, max(1, nb // 64)))[:64]
    for kernel, warps in CONFIGS:
        if time.time() - T0 > budget:
            break
        if cls == "E0" and (kernel, warps) not in ((0, 0), (1, 3), (2, 16)):
            continue
        t = timed(w, kernel, warps)
        lens = w.clen.clone()
        slots = w.slots.view(nb, w.slot)
        bytes_ = [slots[i, : int(lens[i])].clone() for i in sample]
        same = None
        if kernel == 0:
            ref_len, ref_bytes = lens, bytes_
        if ref_len is not None:
            same = bool(torch.equal(lens, ref_len)) and all(torch.equal(a, b) for a, b in zip(bytes_, ref_bytes))
        row = {"class": cls, "blocks": nb, "hc_kernel": kernel, "warps_per_sm": warps, "seconds": round(t, 4),
               "gbs": round(nb * BLOCK / t / GB, 2), "ratio": round(int(lens.sum()) / (nb * BLOCK), 4), "same_as_kernel0": same}
        out["runs"].append(row)
        print(row, flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/hc_ab.json", "w"), indent=1)
    # round trip of the last configuration's output
    for wv in range(w.n_waves):
        b0, b1 = w.decode_wave(wv); torch.cuda.synchronize()
        assert torch.equal(w.out[: (b1 - b0) * BLOCK], w.raw[b0 * BLOCK: b1 * BLOCK])
    del w; torch.cuda.empty_cache()

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/hc_ab.json", "w"), indent=1)
print("done in", round(time.time() - T0, 1), "s", flush=True)
