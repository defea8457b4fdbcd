#!/bin/bash
# round 2, GPU call 2: lane-per-block decoder -- correctness on long runs, ncu --set full on E50 and ETEXT
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== LPB on long-run classes ==="
timeout 300 python tools/sweep.py 16384 E0,E100 dec 32,1 2>&1 | tee gpurun_out/r2c3_lpb_longruns.txt
echo "=== pytest decode ==="
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or roundtrip or large_batch" 2>&1 | tail -8 | tee gpurun_out/r2c3_pytest.txt
for cls in E50 ETEXT; do
  echo "=== ncu LPB $cls ==="
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lpb -s 1 -c 1 -f -o gpurun_out/lpb_${cls}_r2c3 python tools/dec_one.py 42624 $cls 1 > gpurun_out/ncu_lpb_${cls}_r2c3.log 2>&1; tail -2 gpurun_out/ncu_lpb_${cls}_r2c3.log
done
timeout 300 python tools/sweep.py 42624 E50,ETEXT dec 108,104,1,2 2>&1 | tee gpurun_out/r2c3_dec_sweep.txt
