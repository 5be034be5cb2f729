#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run35.log; : > $L
for emu in 2 0 4; do
  echo "=== attention timing EMU=$emu (nomax)" >> $L
  MGLD_ATTN_EMU=$emu timeout 200 python tools/dev_check_attention_v3.py child 2>&1 | grep -E "self B5 N4096 h5 qscale1|self B10 N4096|self B10 N1024|rror" >> $L
done
for mt in 512 64; do
  echo "=== perf unet T=10 + vae, pair_min_tiles=$mt" >> $L
  MGLD_CONV_PAIR_MIN_TILES=$mt MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "graph:|rror" >> $L
  MGLD_CONV_PAIR_MIN_TILES=$mt timeout 300 python tools/dev_perf_e2e.py vae 2>&1 | grep -E "VAE|rror" >> $L
done
cat $L | cut -c1-200
