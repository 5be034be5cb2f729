#!/bin/bash
# programmatic dependent launch: parity (eager + graph), tile-step and clip timing with MGLD_PDL=1 / 0
mkdir -p gpurun_out
L=gpurun_out/run29.log; : > $L
echo "=== pytest ops + models (PDL on)" >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q --timeout=300 -x >> $L 2>&1
echo "exit=$?" >> $L
for pdl in 1 0; do
  echo "=== perf unet T=10 MGLD_PDL=$pdl" >> $L
  MGLD_PDL=$pdl MGLD_T=10 timeout 300 python tools/dev_perf_e2e.py unet 2>&1 | grep -E "eager|graph:|rror|finite" >> $L
done
echo "=== pipeline phases (PDL on)" >> $L
timeout 600 python tools/dev_pipeline_phases.py >> $L 2>&1
grep -E "exit=|eager|graph:|===|rror|passed|failed|clip total|finite" $L | cut -c1-200 | tail -30
