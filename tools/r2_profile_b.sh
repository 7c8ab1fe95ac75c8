#!/bin/bash
# ncu evidence after the second session of round 2 (warp-per-item cross-attention, tanh-form GELU): launch list of one warm round and
# --set full of one whole decode step.  Output: gpurun_out/r2q_* (text summaries).
set -u
mkdir -p gpurun_out
R="python tools/profile_round.py --hist 150"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2q_launches_warm.csv $R > gpurun_out/r2q_ncu0.log 2>&1
python tools/agg_launches.py gpurun_out/r2q_launches_warm.csv > gpurun_out/r2q_launches_warm.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none -s 533 -c 103 -f -o gpurun_out/r2q_decode_step $R > gpurun_out/r2q_decode_step.log 2>&1
python tools/ncu_summary.py gpurun_out/r2q_decode_step.ncu-rep > gpurun_out/r2q_decode_step_summary.txt 2>&1
python tools/ncu_condense.py gpurun_out/r2q_decode_step_summary.txt > gpurun_out/r2q_decode_step_by_kernel.txt 2>&1
rm -f gpurun_out/r2q_decode_step.ncu-rep
R2="python tools/profile_round.py --hist 256"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:dec_cross_warp" -s 24 -c 4 -f -o gpurun_out/r2q_cross_h256 $R2 > gpurun_out/r2q_cross_h256.log 2>&1
python tools/ncu_summary.py gpurun_out/r2q_cross_h256.ncu-rep > gpurun_out/r2q_cross_h256_summary.txt 2>&1
rm -f gpurun_out/r2q_cross_h256.ncu-rep
head -30 gpurun_out/r2q_launches_warm.txt; head -20 gpurun_out/r2q_decode_step_by_kernel.txt | cut -c1-250
