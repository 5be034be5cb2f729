#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run7.log; : > $L
echo "=== pytest gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf vae" >> $L
timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
echo "=== ncu norms" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_stats|gn_apply|layernorm" -s 6 -c 6 -o gpurun_out/prof_norms_r01 -f python tools/ncu_norm_target.py >> $L 2>&1
echo "exit=$?" >> $L
grep -E "passed|failed|exit=|eager|graph:|VAE" $L | tail -30
