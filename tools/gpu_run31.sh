#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run31.log; : > $L
echo "=== pytest attention + models + raft" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -k "attention or sample_canvas or raft or clips" >> $L 2>&1
echo "exit=$?" >> $L
for cfg in "1 0" "0 0" "1 500" "1 1000" "1 1500"; do
  set -- $cfg
  echo "=== timing NOMAX=$1 STAGGER=$2" >> $L
  MGLD_ATTN_NOMAX=$1 MGLD_ATTN_STAGGER=$2 timeout 200 python tools/dev_check_attention_v3.py child 2>&1 | grep -E "self B5 N4096 h5 qscale|self B10 N4096|self B10 N1024|rror" >> $L
done
echo "=== counters (NOMAX=1)" >> $L
timeout 200 python tools/dev_attn_counters.py 2>&1 | head -18 >> $L
echo "=== pipeline phases" >> $L
timeout 600 python tools/dev_pipeline_phases.py >> $L 2>&1
grep -E "exit=|===|rror|passed|failed|self B|clip total|n=|mma\.|sm0\.|sm1\.|kernel" $L | cut -c1-200 | tail -60
