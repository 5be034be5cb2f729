#!/bin/bash
# attention v3 (accuracy + speed of the variants), clip batching parity on the GPU, T=5 vs T=10 tile-step, quick bench
mkdir -p gpurun_out
L=gpurun_out/run19.log; : > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
echo "=== attention v3 variants" >> $L
timeout 400 python tools/dev_check_attention_v3.py >> $L 2>&1
echo "exit=$?" >> $L
echo "=== pytest (attention, clip batching)" >> $L
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "attention or clips or sample_canvas" --timeout=200 >> $L 2>&1
echo "exit=$?" >> $L
for T in 5 10; do
  echo "=== perf unet T=$T" >> $L
  MGLD_T=$T timeout 300 python tools/dev_perf_e2e.py unet >> $L 2>&1
done
echo "=== bench (no cpu baseline)" >> $L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r01d.json 2>> $L
echo "exit=$?" >> $L
cat gpurun_out/bench_r01d.json >> $L
grep -E "exit=|eager|graph:|===|rror|passed|failed|self B|cross B|== v|\"value\"" $L | cut -c1-220 | tail -70
