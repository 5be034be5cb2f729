#!/bin/bash
# A/B: mbarrier try_wait suspend hint (default lib) vs plain polling (build/libmgld_nosuspend.so); new GELU / small-cout /
# gn_apply kernels parity; cross-attention v1 vs v3
mkdir -p gpurun_out
L=gpurun_out/run26.log; : > $L
echo "=== pytest ops + models (suspend build)" >> $L
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
for lib in "" "mgld-vsr_b200/build/libmgld_nosuspend.so"; do
  echo "=== attention timing lib=[$lib]" >> $L
  MGLD_LIB=$lib timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
  echo "=== perf unet T=10 lib=[$lib]" >> $L
  MGLD_LIB=$lib MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet >> $L 2>&1
done
echo "=== attention v1 forced (cross-attention comparison)" >> $L
MGLD_ATTN_V1=1 timeout 200 python tools/dev_check_attention_v3.py child >> $L 2>&1
echo "=== layer table T=10" >> $L
MGLD_T=10 timeout 300 python tools/dev_layer_table.py >> $L 2>&1
grep -E "exit=|eager|graph:|===|rror|passed|failed|self B5 N4096 h5 qscale1|self B10|cross B" $L | cut -c1-200 | tail -50
