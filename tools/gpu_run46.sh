#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run46.log; : > $L
echo "=== pytest ops + models" >> $L
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=200 >> $L 2>&1
echo "exit=$?" >> $L
for r in 1 2; do
echo "=== perf unet T=10 (SPADE epilogue hoist, M-outer tile order) run $r" >> $L
MGLD_T=10 timeout 200 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "graph:|rror|finite" >> $L
done
cat $L | cut -c1-200 | tail -12
