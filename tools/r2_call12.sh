#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=r2c14
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or roundtrip or large_batch" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
timeout 600 python tools/sweep.py 75776 E50,ETEXT dec 108,104,1 2>&1 | tee gpurun_out/${T}_dec_sweep.txt
timeout 600 python tools/sweep.py 262144 E50,ETEXT dec 108,104,1 2>&1 | tee -a gpurun_out/${T}_dec_sweep.txt
timeout 900 python bench.py --blocks 262144 --steps 3 --warmup 3 --no-hc --no-cpu --no-e2e --stream-gib 4 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['decode_gbs']); [print(k, v['decode_gbs'], v['decode_tuned_gbs']) for k,v in d['entropy_sweep'].items()]"; tail -3 gpurun_out/${T}_bench.err
