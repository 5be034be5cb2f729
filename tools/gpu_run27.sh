#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run27.log; : > $L
echo "=== pytest ops + models" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
echo "=== attention timing" >> $L
timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
echo "=== perf unet T=10" >> $L
MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "=== pipeline phases" >> $L
timeout 600 python tools/dev_pipeline_phases.py >> $L 2>&1
grep -E "exit=|eager|graph:|===|rror|passed|failed|self B5 N4096 h5 qscale1|self B10|cross B|clip total|n=" $L | cut -c1-200 | tail -40
