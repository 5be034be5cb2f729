#!/bin/bash
# Multi-GPU evidence runs (run under `gpurun --gpus N`): tools/run_multigpu_checks.sh <N> <config> [extra bench args]
# Launches bench.py exactly as the driver does (torchrun, one rank per GPU, 127.0.0.1 rendezvous).
set -euo pipefail
N=$1; CFG=$2; shift 2
mkdir -p gpurun_out
if [[ "$N" == "1" ]]; then
  python bench.py --gpus 1 --config "$CFG" "$@"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus "$N" --config "$CFG" "$@"
fi
