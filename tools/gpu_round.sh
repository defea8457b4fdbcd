#!/bin/bash
# One gpurun call's worth of measurement: bench (default = BASELINE configs[1]), the reference arm, the ncu launch list
# and ncu --set full of the decode and fast-encode kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m lz4net_b200.build > /dev/null 2>&1
echo "=== bench (default) ==="
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 5000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
echo "=== bench --impl reference ==="
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
# 151552 blocks = two whole waves of the lane-per-block decoder: that kernel decodes every block of the launch
SMALL="--blocks 151552 --steps 2 --warmup 1 --no-sweep --no-hc --no-cpu --no-e2e --no-stream"
echo "=== ncu launch list ==="
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
grep -c lz4 gpurun_out/launches_$TAG.csv
echo "=== ncu full: decode ==="
ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lpb -s 3 -c 1 -f -o gpurun_out/dec_$TAG python bench.py $SMALL > gpurun_out/ncu_dec_$TAG.log 2>&1; tail -1 gpurun_out/ncu_dec_$TAG.log
echo "=== ncu full: encode ==="
ncu --set full --clock-control none --import-source on -k regex:lz4_encode_fast -s 3 -c 1 -f -o gpurun_out/enc_$TAG python bench.py $SMALL > gpurun_out/ncu_enc_$TAG.log 2>&1; tail -1 gpurun_out/ncu_enc_$TAG.log
