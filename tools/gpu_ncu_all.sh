#!/bin/bash
# One ncu session: launch list of one eager denoise step, full capture of the first six tcgen05 launches of a warm step
# (in-proj, QKV, attention, out-proj+LN, FFN1+GELU, FFN2+LN), and a full capture of the update kernel alone at B=2048 (HBM-bound).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --eager > gpurun_out/profile_step.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tc_|update_kernel|token0|motion_to" -s 135 -c 90 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --eager > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_|update_kernel" -s 129 -c 6 -o gpurun_out/prof_step python tools/profile_step.py --eager > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
timeout 120 python tools/profile_update.py > gpurun_out/profile_update.log 2>&1 || { echo "plain update run failed"; tail -20 gpurun_out/profile_update.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:update_kernel" -s 3 -c 2 -o gpurun_out/prof_update python tools/profile_update.py > gpurun_out/ncu_update.log 2>&1
echo "ncu update rc=$?"
python tools/ncu_raw_summary.py gpurun_out/prof_step.ncu-rep > gpurun_out/prof_step_summary.txt 2>&1
python tools/ncu_raw_summary.py gpurun_out/prof_update.ncu-rep > gpurun_out/prof_update_summary.txt 2>&1
du -sh gpurun_out; cat gpurun_out/profile_update.log; tail -12 gpurun_out/profile_step.log
# the finetune step's kernels (eager launches): the tensor-core training GEMM, the fused short-sequence attention, conversions
MST_TRAIN_GRAPH=0 timeout 200 python tools/bench_finetune.py --sg 1 --steps 1 --warmup 1 > gpurun_out/ft_plain.log 2>&1 || { echo "plain finetune run failed"; tail -20 gpurun_out/ft_plain.log; exit 1; }
MST_TRAIN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_kernel|attn_small|layernorm_bwd" -s 40 -c 8 -o gpurun_out/prof_finetune python tools/bench_finetune.py --sg 1 --steps 1 --warmup 1 > gpurun_out/ncu_finetune.log 2>&1
echo "ncu finetune rc=$?"
python tools/ncu_raw_summary.py gpurun_out/prof_finetune.ncu-rep > gpurun_out/prof_finetune_summary.txt 2>&1
du -sh gpurun_out
