#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run6.log; : > $L
for grp in "a2" "a3 a4 a6" "x1" "l1"; do
  echo "=== $grp" >> $L
  timeout 120 python tools/dev_check_attn_flow.py $grp >> $L 2>&1
  echo "exit=$?" >> $L
done
echo "=== perf attention v2" >> $L
timeout 200 python tools/dev_perf_attention.py >> $L 2>&1
echo "=== perf attention v1" >> $L
MGLD_ATTN_V1=1 timeout 200 python tools/dev_perf_attention.py >> $L 2>&1
echo "=== pytest attention" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k attention --timeout=120 >> $L 2>&1
echo "exit=$?" >> $L
grep -E "PASS|FAIL|ERROR|passed|failed|exit=|attn" $L | tail -60
