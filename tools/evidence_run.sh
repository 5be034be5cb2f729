#!/bin/bash
# Round evidence run (1 GPU, under gpurun): GPU tests, smoke, both bench arms, the ncu launch list of one tile-step at the
# bench's batch (T=10) and `ncu --set full` captures of the two tensor-core kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 3000 -- 'bash tools/evidence_run.sh r02'
set -uo pipefail
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
echo "== pytest -m gpu"; python -m pytest tests -m gpu -q -s > $o/${tag}_gpu_tests.log 2>&1; tail -3 $o/${tag}_gpu_tests.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/${tag}_smoke.log
echo "== bench (mgld arm)"; python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; head -c 400 $o/${tag}_bench.json; echo
echo "== bench (reference arm)"; python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err; head -c 300 $o/${tag}_bench_reference_arm.json; echo
if [ "${SKIP_LIST:-0}" != "1" ]; then
echo "== ncu launch list (one eager struct-encoder + UNet tile-step, T=10)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
MGLD_T=10 MGLD_PDL=0 timeout 1500 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $o/${tag}_ncu_launches_tile_step_T10.csv python tools/ncu_target.py 2 > $o/${tag}_ncu_launches.log 2>&1
wc -l $o/${tag}_ncu_launches_tile_step_T10.csv
fi
echo "== ncu --set full: attention (64x64 self-attention, B=5)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_v3 -s 1 -c 1 -f -o $o/${tag}_prof_attention python tools/ncu_attn_target.py > $o/${tag}_ncu_attn.log 2>&1
echo "== ncu --set full: cross-attention against the 77 text tokens (warp-level MMA kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attention_kv80 -s 1 -c 1 -f -o $o/${tag}_prof_attention_kv80 python tools/ncu_attn_target.py > $o/${tag}_ncu_attn_kv80.log 2>&1
echo "== ncu --set full: head-dim-512 flash attention of the VAE middle block (one frame of a 960x960 tile)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_hd512 -c 1 -f -o $o/${tag}_prof_attention_hd512 python tools/ncu_vae_attn_target.py > $o/${tag}_ncu_attn_hd512.log 2>&1
echo "== ncu --set full: conv_gemm (six representative launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 6 -f -o $o/${tag}_prof_conv_gemm python tools/ncu_conv_target.py > $o/${tag}_ncu_conv.log 2>&1
ls -la $o | grep ${tag}_prof
