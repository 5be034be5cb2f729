#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run16.log; : > $L
echo "=== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 1500 python bench.py > gpurun_out/bench_r01c.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01c.json >> $L
echo "=== bench reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r01c_ref.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01c_ref.json >> $L
grep -E "smoke|exit=|metric|rror" $L | cut -c1-1500 | tail -12
