#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run12.log; : > $L
echo "=== pytest ops (pair auto)" >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pytest ops (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x -k "gemm or conv" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== counters" >> $L
timeout 300 python tools/dev_conv_counters.py >> $L 2>&1
echo "=== perf conv (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 300 python tools/dev_perf_conv_gemm.py >> $L 2>&1
echo "=== perf conv (pair on)" >> $L
MGLD_CONV_PAIR=1 timeout 300 python tools/dev_perf_conv_gemm.py >> $L 2>&1
grep -E "passed|failed|exit=|eager|graph:|VAE|TFLOP|split=|===|rror|pair=" $L | cut -c1-250 | tail -120
