#!/bin/bash
# attention v3: single-lane arrives, late pv wait, pipelined score load, warp-uniform MMA issue loop
mkdir -p gpurun_out
L=gpurun_out/run22.log; : > $L
echo "=== role counters" >> $L
timeout 200 python tools/dev_attn_counters.py >> $L 2>&1
for emu in 2 4 0; do
  echo "=== timing EMU=$emu" >> $L
  MGLD_ATTN_EMU=$emu timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
  echo "exit=$?" >> $L
done
echo "=== pytest (attention)" >> $L
timeout 400 python -m pytest tests/test_ops_gpu.py -q -k "attention" --timeout=200 >> $L 2>&1
echo "exit=$?" >> $L
grep -E "exit=|===|rror|self B|cross B|mma\.|sm0\.|sm1\.|tma\.|kernel|CTAs|passed|failed" $L | cut -c1-200 | tail -120
