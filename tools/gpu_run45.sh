#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run45.log; : > $L
echo "=== conv tests" >> $L
timeout 300 python -m pytest tests/test_ops_gpu.py -q --timeout=200 -k "conv_gemm or linear or geglu or spade" >> $L 2>&1
echo "exit=$?" >> $L
for r in 1 2; do
echo "=== perf unet T=10 (MMA issuer + epilogue waits polling, producers parked) run $r" >> $L
MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "graph:|rror" >> $L
done
cat $L | cut -c1-200 | tail -12
