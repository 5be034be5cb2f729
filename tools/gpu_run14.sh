#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run14.log; : > $L
echo "=== pytest ops (pair auto)" >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pytest ops (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x -k "gemm or conv" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== counters" >> $L
timeout 300 python tools/dev_conv_counters.py 0 >> $L 2>&1
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "=== perf unet (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "=== perf vae" >> $L
timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
grep -E "passed|failed|exit=|eager|graph:|VAE|===|rror|pair=" $L | cut -c1-330 | tail -60
