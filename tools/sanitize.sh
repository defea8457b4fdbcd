#!/bin/bash
# compute-sanitizer passes over a small encode + decode (memcheck: no access outside the buffers; racecheck: the
# shared-memory protocols of the encoder rounds and the decoder ring / stage are properly ordered).  Run on a GPU box.
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, oracle, lz4net_b200
from tests import cases
ctx = lz4net_b200.default_context()
blocks = [cases.content(m, n, seed=i).tobytes() for i, (m, n) in enumerate((("E50", 65536), ("ETEXT", 30000), ("E100", 65536), ("periodic", 20000), ("E0", 9000), ("lowent", 65546), ("mixed", 70000)))]
res, outs = ctx.encode_blocks(blocks)
for b, r, o in zip(blocks, res, outs):
    assert (r, o) == oracle.encode(b)
for lanes in (32, 16, 108, 104, 1):                 # 1 = the lane-per-block kernel (round 2)
    ctx.set_option("decode_lanes", lanes)
    for known in (True, False):
        r2, dec = ctx.decode_blocks(outs, [len(b) for b in blocks], known=known)
        assert dec == blocks
# untrusted input: truncated and bit-flipped streams must be rejected or decoded WITHOUT touching memory outside the buffers
rng = np.random.default_rng(3)
bad = []
for o in outs:
    a = bytearray(o)
    bad.append(bytes(a[: max(1, len(a) // 2)]))
    for _ in range(8):
        a[int(rng.integers(0, len(a)))] ^= 1 << int(rng.integers(0, 8))
    bad.append(bytes(a))
for lanes in (1, 16, 104):
    ctx.set_option("decode_lanes", lanes)
    for known in (True, False):
        ctx.decode_blocks(bad, [len(b) for b in blocks for _ in (0, 1)], known=known)
print("case ok")
PY
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "case ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
