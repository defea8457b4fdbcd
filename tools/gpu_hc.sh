#!/bin/bash
# One short gpurun call for the LZ4HC kernels: oracle parity of every kernel and of the per-batch choice (pytest subset),
# smoke(), the kernels side by side (tools/hc_ab.py).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
echo "=== pytest HC subset ==="
timeout -s KILL 140 python -m pytest tests/test_gpu_parity.py -q --timeout 80 -k "hc or golden or wrap or lz4stream or single_block or limited_output" 2>&1 | tail -8 | tee gpurun_out/pytest_hc.txt
echo "elapsed $(( $(date +%s) - T0 )) s"
echo "=== smoke ==="
timeout -s KILL 45 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "elapsed $(( $(date +%s) - T0 )) s"
echo "=== hc_ab ==="
timeout -s KILL 120 python tools/hc_ab.py 85 2>&1 | tail -40 | tee gpurun_out/hc_ab.txt
echo "elapsed $(( $(date +%s) - T0 )) s"
