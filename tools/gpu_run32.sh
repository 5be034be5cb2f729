#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run32.log; : > $L
MGLD_T=10 timeout 300 python tools/dev_conv_counters.py 0 >> $L 2>&1
cat $L | cut -c1-600
