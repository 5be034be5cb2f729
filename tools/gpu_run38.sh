#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run38.log; : > $L
echo "=== pytest ops + models" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet T=10" >> $L
MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "eager|graph:|rror|finite|conv_gemm" | cut -c1-62,150-250 >> $L
echo "=== pipeline phases" >> $L
timeout 600 python tools/dev_pipeline_phases.py 2>&1 | grep -E "clip total|n=" >> $L
cat $L | cut -c1-200 | tail -28
