#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c6
echo "=== pytest encode ==="
timeout 900 python -m pytest tests -m gpu -x -q -k "encode or large_batch" 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.txt
echo "=== encoder sweep: tagged lane-per-block warp off / on ==="
timeout 900 python tools/enc_sweep.py 262144 E50,ETEXT 0 512 2 0,1 2>&1 | tee gpurun_out/${T}_enc_sweep.txt
timeout 300 python tools/enc_sweep.py 131072 E0,E100 0 512 2 0,1 2>&1 | tee -a gpurun_out/${T}_enc_sweep.txt
echo "=== ncu encode with lane warp E50 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_fast -s 1 -c 1 -f -o gpurun_out/enc_lw_E50_${T} python tools/enc_one.py 131072 E50 2 512 0 1 > gpurun_out/ncu_enc_lw_E50_${T}.log 2>&1; tail -1 gpurun_out/ncu_enc_lw_E50_${T}.log
