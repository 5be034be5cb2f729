#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run11.log; : > $L
echo "=== pytest ops (pair auto)" >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pytest ops (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout=300 -x -k "gemm or conv" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf conv (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 300 python tools/dev_perf_conv_gemm.py >> $L 2>&1
echo "=== ncu conv" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 6 -f -o gpurun_out/prof_conv_pair_r01 python tools/ncu_conv_target.py >> $L 2>&1
grep -E "passed|failed|exit=|eager|graph:|VAE|TFLOP|split=|===|rror" $L | cut -c1-200 | tail -90
