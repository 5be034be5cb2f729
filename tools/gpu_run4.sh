#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run4.log; : > $L
for grp in "g1" "g2 g3 g4 g5 g6 g7 g8" "c1 c2 c3 c4 c5 c6" "t1 t2 e1 e2 s1 s2"; do
  echo "=== $grp" >> $L
  timeout 120 python tools/dev_check_conv_gemm.py $grp >> $L 2>&1
  echo "exit=$?" >> $L
done
echo "=== pytest gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf conv" >> $L
timeout 300 python tools/dev_perf_conv_gemm.py >> $L 2>&1
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf vae" >> $L
timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
echo "exit=$?" >> $L
grep -E "PASS|FAIL|ERROR|passed|failed|exit=|eager|graph|VAE|TFLOP" $L | tail -120
