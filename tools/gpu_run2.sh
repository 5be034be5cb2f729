#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run2.log; : > $L
for grp in "a1" "a2 a3 a4 a5 a6" "x1 x2" "l1" "l2 l3" "fl" "gd"; do
  echo "=== $grp" >> $L
  timeout 120 python tools/dev_check_attn_flow.py $grp >> $L 2>&1
  echo "exit=$?" >> $L
done
echo "=== perf" >> $L
timeout 200 python tools/dev_perf_conv_gemm.py >> $L 2>&1
tail -120 $L
