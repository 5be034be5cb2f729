#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run17.log; : > $L
echo "=== ncu --set full, all conv_gemm launches of one eager tile-step" >> $L
timeout 1500 ncu --set full --clock-control none --import-source off -k regex:conv_gemm_kernel -f -o gpurun_out/prof_conv_tilestep_r01 python tools/ncu_target.py 1 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== ncu --set full, attention + norm kernels (first 24)" >> $L
timeout 900 ncu --set full --clock-control none --import-source off -k regex:"attention|gn_|layernorm" -c 24 -f -o gpurun_out/prof_attn_norm_r01 python tools/ncu_target.py 1 >> $L 2>&1
echo "exit=$?" >> $L
ls -la gpurun_out/*.ncu-rep >> $L
tail -5 $L
