#!/bin/bash
# round 2, GPU call 1: parity tests, the encoder warps-per-SM sweep (global-table warps) and the decoder sweep incl. the
# lane-per-block kernel.  Every step under its own timeout; logs in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/r2c1_gpu.txt
echo "=== lane-per-block decoder quick check ==="
timeout 300 python tools/sweep.py 16384 E50,ETEXT dec 108,1,2 2>&1 | tee gpurun_out/r2c1_lpb_quick.txt
echo "=== pytest -m gpu ==="
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2c1_pytest.txt
echo "=== decoder sweep ==="
timeout 600 python tools/sweep.py 65536 E50,ETEXT dec 108,104,8,4,1,2 2>&1 | tee gpurun_out/r2c1_dec_sweep.txt
timeout 300 python tools/sweep.py 32768 E0,E100 dec 32,16,1 2>&1 | tee -a gpurun_out/r2c1_dec_sweep.txt
echo "=== encoder sweep: warps per SM x global-table variant ==="
timeout 900 python tools/enc_sweep.py 65536 E50,ETEXT 14,18,20,24,28 512 2,12 2>&1 | tee gpurun_out/r2c1_enc_sweep.txt
timeout 300 python tools/enc_sweep.py 32768 E0,E100 14,20,28 512 2 2>&1 | tee -a gpurun_out/r2c1_enc_sweep.txt
