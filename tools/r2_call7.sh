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
echo "=== bench e2e only (small headline) ==="
timeout 900 python bench.py --blocks 65536 --steps 2 --warmup 3 --no-sweep --no-hc --no-cpu --no-stream > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]); print({k:v for k,v in d['e2e'].items() if k!='sample'})"; tail -3 gpurun_out/${T}_bench.err
