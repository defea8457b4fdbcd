#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c7
echo "=== pytest ==="
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.txt
echo "=== encoder sweep ==="
timeout 900 python tools/enc_sweep.py 131072 E50,ETEXT,E0,E100 0 512 2 2>&1 | tee gpurun_out/${T}_enc_sweep.txt
echo "=== decoder sweep (LPB geometry B: 17 warps) ==="
timeout 600 python tools/sweep.py 80512 E50,ETEXT dec 108,104,1,2 2>&1 | tee gpurun_out/${T}_dec_sweep.txt
