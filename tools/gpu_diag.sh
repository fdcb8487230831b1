#!/bin/bash
# Run every GPU test function in its own process (a device trap in one kernel must not poison the
# CUDA context of the others) and write one log per function under gpurun_out/diag/.
# usage (on the GPU box):  bash tools/gpu_diag.sh [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/diag
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/diag/_gpu.txt 2>&1
sel=${1:-}
funcs=$(python -m pytest tests -m gpu --collect-only -q ${sel:+-k "$sel"} 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u)
: > gpurun_out/diag/_summary.txt
for f in $funcs; do
  name=$(echo "$f" | sed 's#.*::##')
  timeout 300 python -m pytest "$f" -q --timeout 120 -p no:cacheprovider > "gpurun_out/diag/$name.log" 2>&1
  rc=$?
  echo "$rc $f $(tail -1 gpurun_out/diag/$name.log)" | tee -a gpurun_out/diag/_summary.txt
done
