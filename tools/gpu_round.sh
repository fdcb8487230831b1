#!/bin/bash
# One GPU session: full GPU test-suite, smoke, bench, then the ncu launch list and one full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" != "noncu" ]; then
timeout 300 python tools/profile_step.py --eager > gpurun_out/profile_step.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 90 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --eager > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 35 -c 4 -o gpurun_out/prof_gemm python tools/profile_step.py --eager > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
fi
[ -f gpurun_out/profile_step.log ] && cat gpurun_out/profile_step.log
exit 0
