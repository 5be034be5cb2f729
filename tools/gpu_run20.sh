#!/bin/bash
# attention v3 with the lean MMA issue loop: role cycle counters + timing; SPADE shared-conv batching parity + tile-step timing
mkdir -p gpurun_out
L=gpurun_out/run20.log; : > $L
echo "=== attention v3 role counters" >> $L
timeout 200 python tools/dev_attn_counters.py >> $L 2>&1
echo "exit=$?" >> $L
echo "=== attention timing (default variant)" >> $L
timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pytest (attention, unet)" >> $L
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "attention or unet or sample_canvas" --timeout=200 >> $L 2>&1
echo "exit=$?" >> $L
for T in 5 10; do
  echo "=== perf unet T=$T" >> $L
  MGLD_T=$T timeout 300 python tools/dev_perf_e2e.py unet >> $L 2>&1
done
grep -E "exit=|eager|graph:|===|rror|passed|failed|self B|cross B|mma\.|sm0\.|sm1\.|tma\.|kernel|CTAs" $L | cut -c1-200 | tail -90
