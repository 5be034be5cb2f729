#!/bin/bash
# first GPU contact: conv_gemm correctness in separate processes (a trap in one must not take the others down)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/run1.log 2>&1
for grp in "g1" "g2 g3 g4" "g5" "g6 g7 g8" "c1 c2 c3" "c4" "c5 c6" "t1 t2" "e1 e2" "s1 s2"; do
  echo "=== $grp" >> gpurun_out/run1.log
  timeout 120 python tools/dev_check_conv_gemm.py $grp >> gpurun_out/run1.log 2>&1
  echo "exit=$?" >> gpurun_out/run1.log
done
tail -60 gpurun_out/run1.log
