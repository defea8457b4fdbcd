#!/bin/bash
# One short gpurun call for the LZ4HC kernels: oracle parity of every kernel (pytest subset), the kernels side by side
# (tools/hc_ab.py), then -- if time is left -- an ncu capture of the warp kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
echo "=== pytest HC subset ==="
timeout -s KILL 170 python -m pytest tests/test_gpu_parity.py -q --timeout 80 -k "hc_byte_identical or hc_limited_output_every or hc_handed_back or device_batch_hc" 2>&1 | tail -8 | tee gpurun_out/pytest_hc.txt
echo "elapsed $(( $(date +%s) - T0 )) s"
echo "=== hc_ab ==="
timeout -s KILL 170 python tools/hc_ab.py 120 2>&1 | tail -30 | tee gpurun_out/hc_ab.txt
echo "elapsed $(( $(date +%s) - T0 )) s"
echo "=== ncu: warp kernels, one small batch each ==="
timeout -s KILL 100 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_hcw -c 2 -f -o gpurun_out/hcw_r02c python tools/hc_one.py > gpurun_out/ncu_hcw.log 2>&1; tail -2 gpurun_out/ncu_hcw.log
echo "elapsed $(( $(date +%s) - T0 )) s"
