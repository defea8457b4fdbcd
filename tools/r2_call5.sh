#!/bin/bash
# round 2, GPU call 5: the new bench.py (defaults, e2e variants, sweep, stream) at N=1 with a reduced headline, full pytest
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c5
echo "=== pytest -m gpu ==="
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.txt
echo "=== bench (reduced: 262144 blocks) ==="
timeout 1500 python bench.py --blocks 262144 --steps 3 --warmup 3 --stream-gib 4 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 6000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
