#!/bin/bash
# new conv_small_cin / layernorm kernels (parity), per-layer table of the T=10 tile-step, tile-step timing
mkdir -p gpurun_out
L=gpurun_out/run24.log; : > $L
echo "=== pytest ops" >> $L
timeout 500 python -m pytest tests/test_ops_gpu.py -q --timeout=200 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== layer table T=10" >> $L
MGLD_T=10 timeout 300 python tools/dev_layer_table.py >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet T=10" >> $L
MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet >> $L 2>&1
grep -E "exit=|eager|graph:|===|rror|passed|failed" $L | cut -c1-200 | tail -20
