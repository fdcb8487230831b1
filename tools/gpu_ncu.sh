#!/bin/bash
# ncu: launch list of two eager steps + full capture of selected kernels (regex $1, skip $2, count $3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --eager > gpurun_out/profile_step.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 138 -c 90 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --eager > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${1:-tc_gemm}" -s ${2:-35} -c ${3:-4} -o gpurun_out/prof_${4:-gemm} python tools/profile_step.py --eager > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
tail -12 gpurun_out/profile_step.log
