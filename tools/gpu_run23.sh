#!/bin/bash
# ncu --set full with source-level warp-state sampling of ONE attention_v3 launch (UNet 64x64 self-attention shape)
mkdir -p gpurun_out
L=gpurun_out/run23.log; : > $L
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_v3 -s 2 -c 1 -f -o gpurun_out/prof_attention_v3_r01 python tools/ncu_attn_target.py >> $L 2>&1
echo "exit=$?" >> $L
ls -la gpurun_out/*.ncu-rep >> $L
tail -5 $L
