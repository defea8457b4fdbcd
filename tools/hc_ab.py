"""LZ4HC kernels side by side on one GPU: the default (kernel chosen per batch, hc_kernel -1), the thread-per-block kernel
(0) and the warp-per-block kernels on a static index (1: block staged in shared memory, 2: block through L1) at several
residencies.  Device-timed, one warm-up + the better of two timed passes per configuration; every configuration's lengths
and a sample of its bytes are compared with kernel 0's (oracle parity is tests/test_gpu_parity.py's job).  Classes: the
synthetic E50 / ETEXT of the bench and TEXT = 64 KiB blocks cut from this repository's own documents (natural text,
average hash-bucket size ~20-50).  Writes gpurun_out/hc_ab.json.  usage: hc_ab.py [seconds]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lz4net_b200
from bench import Workload, BLOCK, GB

T0 = time.time()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0          # seconds this tool may take
ctx = lz4net_b200.Context(0)
out = {"device": torch.cuda.get_device_name(0), "block": BLOCK, "runs": []}
CONFIGS = [(0, 0), (-1, 0), (2, 32), (2, 24), (1, 3), (2, 16), (2, 28), (2, 20)]
if len(sys.argv) > 2:                                                  # e.g. "2:32,2:24,-1:0": a chosen subset (no kernel 0: no byte comparison)
    CONFIGS = [tuple(int(x) for x in c.split(":")) for c in sys.argv[2].split(",")]
CLASSES = (("E50", 65536), ("ETEXT", 32768), ("TEXT", 16384))
LONE = len(sys.argv) <= 2


def timed(w, kernel, warps):
    ctx.set_option("hc_kernel", kernel); ctx.set_option("hc_warps_per_sm", warps)
    w.clen.zero_()
    w.encode(hc=True); torch.cuda.synchronize()
    best = 1e9
    for _ in range(1 if kernel == 0 else 2):          # (the thread kernel takes seconds on text: one pass)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); w.encode(hc=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def text_blocks(nb):
    data = b""
    for name in ("SURVEY.md", "DESIGN.md", "BASELINE.md", "INTEGRATION.md", "README.md", "bench.py", "lz4net_b200/csrc/capi.cu"):
        try:
            data += open(os.path.join(ROOT, name), "rb").read()
        except OSError:
            pass
    k = max(1, len(data) // BLOCK)
    t = torch.frombuffer(bytearray(data[: k * BLOCK]), dtype=torch.uint8).cuda().view(k, BLOCK)
    return t.repeat((nb + k - 1) // k, 1)[:nb].contiguous().view(-1), k


for cls, nb in CLASSES:
    w = Workload(ctx, nb, "E0" if cls == "TEXT" else cls, nb, seed=3)
    if cls == "TEXT":
        w.raw, distinct = text_blocks(nb)
        out["text_distinct_blocks"] = distinct
    ref_len = ref_bytes = None
    sample = list(range(0, nb, max(1, nb // 64)))[:64]
    for kernel, warps in CONFIGS:
        if time.time() - T0 > budget:
            break
        t = timed(w, kernel, warps)
        lens = w.clen.clone()
        slots = w.slots.view(nb, w.slot)
        bytes_ = [slots[i, : int(lens[i])].clone() for i in sample]
        if kernel == 0:
            ref_len, ref_bytes = lens, bytes_
        same = None if ref_len is None else bool(torch.equal(lens, ref_len)) and all(torch.equal(a, b) for a, b in zip(bytes_, ref_bytes))
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

# latency of a lone block (the single-block EncodeHC entry point's case): host timed, one block per call
w = Workload(ctx, 1, "ETEXT", 1, seed=3)
for kernel in ((-1, 0) if LONE else ()):
    ctx.set_option("hc_kernel", kernel); ctx.set_option("hc_warps_per_sm", 0)
    w.encode(hc=True); torch.cuda.synchronize()
    t0 = time.time(); w.encode(hc=True); torch.cuda.synchronize(); dt = time.time() - t0
    out["runs"].append({"class": "ETEXT", "blocks": 1, "hc_kernel": kernel, "seconds": round(dt, 5)})
    print(out["runs"][-1], flush=True)
json.dump(out, open("gpurun_out/hc_ab.json", "w"), indent=1)
print("done in", round(time.time() - T0, 1), "s", flush=True)
