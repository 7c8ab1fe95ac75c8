#!/bin/bash
# ncu evidence for the round-2 default kernels (one gpurun call, one GPU).  Output: gpurun_out/r2p_*.  Numbers printed under a
# profiler are never bench values; per-launch times are serialised (no PDL overlap) - compare shares and utilisation.
set -u
mkdir -p gpurun_out
R="python tools/profile_round.py --hist 150"
# 1. every launch of one warm round, caches left warm between launches
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2p_launches_warm.csv $R > gpurun_out/r2p_ncu0.log 2>&1
python tools/agg_launches.py gpurun_out/r2p_launches_warm.csv > gpurun_out/r2p_launches_warm.txt
# 2. --set full captures
cap() {  # name regex skip count
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c "$4" -f \
      -o "gpurun_out/r2p_$1" $R > "gpurun_out/r2p_$1.log" 2>&1
  python tools/ncu_summary.py "gpurun_out/r2p_$1.ncu-rep" > "gpurun_out/r2p_$1_summary.txt" 2>&1
}
cap enc_gemm_single 'gemm_tc_kernel' 4 6
cap enc_gemm_pair 'gemm_tc2_kernel' 8 4
cap decode_step_gemms 'gemm_tc_kernel' 400 74
cap self_attn 'dec_self_attn_v2' 30 1
cap anc_update 'anc_update' 5 1
cap cross_tma 'dec_cross_tma2' 30 1
cap row_select 'row_select_cluster' 5 1
cap ln_apply 'ln_apply_stats' 5 1
cap enc_attention 'attention_mma' 3 3
cap enc_ln 'add_layernorm_stream' 3 1
ls -la gpurun_out/r2p_* | head -40
