#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run28.log; : > $L
for mt in 512 256 128 64; do
  echo "=== perf unet T=10 pair_min_tiles=$mt" >> $L
  MGLD_CONV_PAIR_MIN_TILES=$mt MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "eager|graph:|conv_gemm|rror" | cut -c1-62,150-250 >> $L
done
cat $L | tail -40
