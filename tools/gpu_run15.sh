#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run15.log; : > $L
echo "=== pytest gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "=== perf vae" >> $L
timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
echo "=== perf vae (pair off)" >> $L
MGLD_CONV_PAIR=0 timeout 600 python tools/dev_perf_e2e.py vae >> $L 2>&1
echo "=== ncu launch list (1 eager tile-step)" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01d.csv python tools/ncu_target.py 1 >> $L 2>&1
grep -E "passed|failed|exit=|eager|graph:|VAE|===|rror|FAILED" $L | cut -c1-250 | tail -40
