#!/bin/bash
# One parameterised GPU lease runner (replaces the per-run scripts of round 1):
#   tools/gpu_run.sh <tag> [--gpus N] [--timeout S] -- <command to run on the box>
# Output of the command goes to gpurun_out/<tag>.log on the box (merged back), gpurun's own verdict to /tmp/<tag>.gpurun.
set -euo pipefail
tag=$1; shift
gpus=1; timeout=1500
while [[ $# -gt 0 && "$1" != "--" ]]; do
  case "$1" in
    --gpus) gpus=$2; shift 2;;
    --timeout) timeout=$2; shift 2;;
    *) echo "unknown option $1" >&2; exit 2;;
  esac
done
shift
extra=(); [[ $gpus -gt 1 ]] && extra=(--gpus "$gpus")
exec /usr/local/graft/bin/gpurun "${extra[@]}" --timeout "$timeout" -- "mkdir -p gpurun_out; ( $* ) > gpurun_out/${tag}.log 2>&1; tail -25 gpurun_out/${tag}.log"
