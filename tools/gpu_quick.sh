#!/bin/bash
# Quick GPU iteration: selected tests (own process each) + the per-kernel event table.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_diag.sh "${1:-tc_gemm or forward or trajectory}" 
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"; cat gpurun_out/profile_step.log | tail -15
