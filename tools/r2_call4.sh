#!/bin/bash
# round 2, GPU call 4: lane-per-block decoder v2.1 and the lane-per-block encoder warp
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c4
echo "=== pytest (encode + decode) ==="
timeout 1200 python -m pytest tests -m gpu -x -q -k "decode or encode or roundtrip or large_batch" 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.txt
echo "=== decoder sweep ==="
timeout 300 python tools/sweep.py 42624 E50,ETEXT dec 108,104,1,2 2>&1 | tee gpurun_out/${T}_dec_sweep.txt
echo "=== encoder sweep: lane-per-block warp off / on ==="
timeout 900 python tools/enc_sweep.py 131072 E50,ETEXT 0 512 2 0,2 2>&1 | tee gpurun_out/${T}_enc_sweep.txt
timeout 300 python tools/enc_sweep.py 32768 E0,E100 0 512 2 0,2 2>&1 | tee -a gpurun_out/${T}_enc_sweep.txt
echo "=== ncu LPB decode E50 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lpb -s 1 -c 1 -f -o gpurun_out/lpb_E50_${T} python tools/dec_one.py 42624 E50 1 > gpurun_out/ncu_lpb_E50_${T}.log 2>&1; tail -1 gpurun_out/ncu_lpb_E50_${T}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lpb -s 1 -c 1 -f -o gpurun_out/lpb_ETEXT_${T} python tools/dec_one.py 42624 ETEXT 1 > gpurun_out/ncu_lpb_ETEXT_${T}.log 2>&1; tail -1 gpurun_out/ncu_lpb_ETEXT_${T}.log
echo "=== ncu encode with lane warp E50 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_fast -s 1 -c 1 -f -o gpurun_out/enc_lw_E50_${T} python tools/enc_one.py 65536 E50 2 512 0 2 > gpurun_out/ncu_enc_lw_E50_${T}.log 2>&1; tail -1 gpurun_out/ncu_enc_lw_E50_${T}.log
