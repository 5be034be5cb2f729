#!/bin/bash
# round-1 evidence run: full GPU test-suite, smoke, bench (both arms), ncu launch list + per-kernel DRAM/tensor metrics of the
# T=10 tile-step, ncu --set full of the attention kernel and of six conv_gemm launches
mkdir -p gpurun_out
L=gpurun_out/run30.log; : > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
echo "=== pytest -m gpu" >> $L
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 1500 python bench.py > gpurun_out/bench_r01e.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01e.json >> $L
echo "=== bench reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r01e_ref.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01e_ref.json >> $L
echo "=== ncu launch list + dram/tensor metrics (1 eager tile-step, T=10)" >> $L
MGLD_T=10 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r01_ncu_launches_tile_step_T10.csv python tools/ncu_target.py 1 >> $L 2>&1
echo "exit=$?" >> $L
echo "=== ncu --set full attention v3 (one launch, with source)" >> $L
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_v3 -s 2 -c 1 -f -o gpurun_out/prof_attention_v3_final_r01 python tools/ncu_attn_target.py >> $L 2>&1
echo "exit=$?" >> $L
echo "=== ncu --set full conv_gemm (six launches)" >> $L
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 6 -f -o gpurun_out/prof_conv_gemm_final_r01 python tools/ncu_conv_target.py >> $L 2>&1
echo "exit=$?" >> $L
ls -la gpurun_out >> $L
grep -E "exit=|===|rror|passed|failed|smoke ok|\"value\"" $L | cut -c1-300 | tail -30
