#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run42.log; : > $L
echo "=== pytest models" >> $L
timeout 600 python -m pytest tests/test_models_gpu.py -q --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pipeline phases (pipelined struct encoder)" >> $L
timeout 600 python tools/dev_pipeline_phases.py 2>&1 | grep -E "clip total|sample_canvas|rror|Traceback" >> $L
cat $L | cut -c1-220 | tail -30
