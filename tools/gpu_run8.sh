#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run8.log; : > $L
echo "=== pytest gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== perf unet" >> $L
timeout 600 python tools/dev_perf_e2e.py unet >> $L 2>&1
echo "=== bench" >> $L
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01b.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01b.json >> $L
grep -E "passed|failed|exit=|eager|graph:|metric" $L | cut -c1-600 | tail -30
