#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c15
timeout 900 python -m pytest tests -m gpu -x -q -k "encode or large_batch" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
timeout 900 python tools/enc_sweep.py 131072 ETEXT,E50,E0,E100 0 512 2 2>&1 | tee gpurun_out/${T}_enc_sweep.txt
