#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run36.log; : > $L
for cfg in "64 18" "16 18" "4 18" "64 10" "64 5" "16 10"; do
  set -- $cfg
  echo "=== perf unet T=10 + vae, pair_min_tiles=$1 min_k=$2" >> $L
  MGLD_CONV_PAIR_MIN_TILES=$1 MGLD_CONV_PAIR_MIN_K=$2 MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "graph:|rror" >> $L
  MGLD_CONV_PAIR_MIN_TILES=$1 MGLD_CONV_PAIR_MIN_K=$2 timeout 300 python tools/dev_perf_e2e.py vae 2>&1 | grep -E "VAE|rror" >> $L
done
cat $L | cut -c1-200
