#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run3.log; : > $L
echo "=== pytest gpu" >> $L
timeout 900 python -m pytest tests -q -m gpu -x --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf vae" >> $L
timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
echo "exit=$?" >> $L
tail -150 $L
