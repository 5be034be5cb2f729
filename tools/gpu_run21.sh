#!/bin/bash
# attention v3: exponential turn-taking between the softmax warpgroups (MGLD_ATTN_SEQ) x exp2 emulation fraction
mkdir -p gpurun_out
L=gpurun_out/run21.log; : > $L
for seq in 1 0; do
  echo "=== role counters SEQ=$seq" >> $L
  MGLD_ATTN_SEQ=$seq timeout 200 python tools/dev_attn_counters.py >> $L 2>&1
done
for cfg in "1 2" "1 4" "1 0" "0 2"; do
  set -- $cfg
  echo "=== timing SEQ=$1 EMU=$2" >> $L
  MGLD_ATTN_SEQ=$1 MGLD_ATTN_EMU=$2 timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
  echo "exit=$?" >> $L
done
grep -E "exit=|===|rror|self B|cross B|mma\.|sm0\.|sm1\.|tma\.|kernel|CTAs" $L | cut -c1-200 | tail -120
