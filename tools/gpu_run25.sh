#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run25.log; : > $L
echo "=== pipeline phases" >> $L
timeout 600 python tools/dev_pipeline_phases.py >> $L 2>&1
echo "exit=$?" >> $L
tail -16 $L
