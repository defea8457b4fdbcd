#!/bin/bash
# The last call of a round: the GPU test suite, the default bench line, the reference arm and the ncu launch list of a
# reduced bench run.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
echo "=== pytest -m gpu ==="
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.txt
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench (default) ==="
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json | head -c 1500; tail -3 gpurun_out/bench_$TAG.err
echo "=== bench --impl reference ==="
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-400 gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
SMALL="--blocks 151552 --steps 2 --warmup 1 --no-sweep --no-hc --no-cpu --no-e2e --no-stream"
echo "=== ncu launch list ==="
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
grep -c lz4 gpurun_out/launches_$TAG.csv
echo "=== LZ4HC kernels side by side ==="
timeout 300 python tools/hc_ab.py 200 2>&1 | tail -40 | tee gpurun_out/hc_ab_$TAG.txt
