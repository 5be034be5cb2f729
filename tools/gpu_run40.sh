#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run40.log; : > $L
echo "=== pytest attention + models" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -k "attention or sample_canvas or unet" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== attention timing" >> $L
timeout 200 python tools/dev_check_attention_v3.py child 2>&1 | grep -E "self B|cross B|rror" >> $L
echo "=== counters" >> $L
timeout 200 python tools/dev_attn_counters.py 2>&1 | head -18 >> $L
cat $L | cut -c1-200 | tail -40
