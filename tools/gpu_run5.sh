#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run5.log; : > $L
echo "=== pytest gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01.json >> $L
echo "=== ncu launches" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01.csv python tools/ncu_target.py 2 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== ncu full conv_gemm" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 300 -c 4 -o gpurun_out/prof_conv_gemm_r01 -f python tools/ncu_target.py 2 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== ncu full attention" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 40 -c 2 -o gpurun_out/prof_attention_r01 -f python tools/ncu_target.py 2 >> $L 2>&1
echo "exit=$?" >> $L
grep -E "passed|failed|exit=|smoke|metric" $L | tail -30
