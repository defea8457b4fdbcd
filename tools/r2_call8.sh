#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c10
echo "=== pytest decode ==="
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or roundtrip or large_batch" 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.txt
echo "=== decoder sweep ==="
timeout 600 python tools/sweep.py 75776 E50,ETEXT dec 108,104,1,2 2>&1 | tee gpurun_out/${T}_dec_sweep.txt
for cls in E50 ETEXT; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lpb -s 1 -c 1 -f -o gpurun_out/lpb2_${cls}_${T} python tools/dec_one.py 75776 $cls 2 > gpurun_out/ncu_lpb2_${cls}_${T}.log 2>&1; tail -1 gpurun_out/ncu_lpb2_${cls}_${T}.log
done
