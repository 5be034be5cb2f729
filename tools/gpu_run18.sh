#!/bin/bash
# round-1 session-2 run: segment batching estimate (T=5 vs T=10 tile-step), ncu --set full of every conv_gemm launch of one
# eager tile-step (CSV export) and of the attention kernels (report with source).
mkdir -p gpurun_out
L=gpurun_out/run18.log; : > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
for T in 5 10; do
  echo "=== perf unet T=$T" >> $L
  MGLD_T=$T timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
done
echo "=== ncu --set full, conv_gemm launches of one eager tile-step" >> $L
timeout 1200 ncu --set full --clock-control none --import-source off -k regex:conv_gemm_kernel -f -o /tmp/prof_conv python tools/ncu_target.py 1 >> $L 2>&1
echo "exit=$?" >> $L
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/r01_ncu_conv_gemm_tilestep_raw.csv 2>> $L
ls -la /tmp/prof_conv.ncu-rep >> $L
sz=$(stat -c %s /tmp/prof_conv.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 30000000 ]; then cp /tmp/prof_conv.ncu-rep gpurun_out/prof_conv_tilestep_r01.ncu-rep; fi
echo "=== ncu --set full, attention kernels" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -c 40 -f -o /tmp/prof_attn python tools/ncu_target.py 1 >> $L 2>&1
echo "exit=$?" >> $L
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/r01_ncu_attention_tilestep_raw.csv 2>> $L
ls -la /tmp/prof_attn.ncu-rep >> $L
sz=$(stat -c %s /tmp/prof_attn.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 25000000 ]; then cp /tmp/prof_attn.ncu-rep gpurun_out/prof_attention_r01.ncu-rep; fi
grep -E "exit=|eager|graph:|===|rror" $L | cut -c1-250 | tail -30
